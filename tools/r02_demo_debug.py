#!/usr/bin/env python
"""Where does the demo-scene render deviate from the reference?  Same features for both, stage by stage."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from oracle import reference_shim as RS, synth
from oracle import render_oracle as RO
ref = RS.install_shim()
from datasets import datas_dict
from models.matchnerf import MatchNeRF as RefNet
from matchnerf_b200.matchnerf import MatchNeRF
from matchnerf_b200 import capi
DEV = "cuda:0"
S = int(os.environ.get("S", "128"))
ds = datas_dict["colmap"](os.path.join(ref, "docs/demo_data"), "test", n_views=3, img_wh=[256, 160], max_len=-1, scene_list=["printer"], test_views_method="fixed", nf_mode="minmax")
batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False)))
enc_sd, dec_sd = synth.synthetic_encoder(1), synth.synthetic_decoder(0)
def rmsf(a, b): return float(((a.double() - b.double()) ** 2).mean().sqrt())
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
opt = RS.reference_options(S, DEV, ref, **{"nerf.rand_rays_test": 20480})
rnet = RefNet(opt).eval(); rnet.feat_enc.load_state_dict(enc_sd); rnet.nerf_dec.load_state_dict(dec_sd); rnet.to(DEV)
onet = MatchNeRF(opt).eval(); onet.feat_enc.load_state_dict(enc_sd); onet.nerf_dec.load_state_dict(dec_sd); onet.to(DEV)
var = RS.EasyDict({k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items()})
print("near_fars", batch["near_fars"][0].tolist())
with torch.no_grad():
    imgs = var.images[:, :3]
    rf = rnet.get_img_feat(imgs)                     # reference features (fp32)
    of = onet.get_img_feat(imgs)
    print("feature rel rms", [rmsf(a, b) / float(b.std()) for a, b in zip(of, rf)])
    tgt, refp = rnet.extract_poses(var)
    H, W = 160, 256
    idx = torch.arange(H * W, device=DEV)
    r = rnet.render(opt, tgt, ray_idx=idx, mode="test", ref_poses=refp, ref_images=imgs, ref_feats_list=rf)
    # ours on the REFERENCE's features, both decoder kernels, both gather kernels
    for gimpl in ("0", "3"):
        os.environ["MNF_GATHER_IMPL"] = gimpl
        ctx = onet.nerf_dec.sync_to_library()
        cpu = lambda t: t.detach().float().cpu()
        packed = ctx.pack_scene([rf[0][0], rf[1][0]], imgs[0], cpu(refp["extrinsics"][0]), cpu(refp["intrinsics"][0]), cpu(refp["near_fars"][0]))
        sc = packed.c_scene(cpu(tgt["extrinsics"][0]), cpu(tgt["intrinsics"][0]), cpu(tgt["near_fars"][0]))
        cfg = onet.nerf_dec.decoder_cfg(opt)
        for impl in (1, 2):
            for mode, kw in (("range", dict(first_ray=0, n_rays=H * W)), ("list", dict(ray_idx=idx))):
                o = ctx.render_rays(sc, cfg, impl=impl, **kw)
                e = (o[0] - r.rgb[0]).abs().amax(-1).view(H, W)
                print(f"gather env {gimpl} decoder impl {impl} rays as {mode}: rgb rms {rmsf(o[0], r.rgb[0]):.2e} opacity rms {rmsf(o[2], r.opacity[0, :, 0]):.2e}; "
                      f"max err {float(e.max()):.3f}; rows with err>0.01: {int((e.amax(1) > 0.01).sum())}, cols: {int((e.amax(0) > 0.01).sum())}")
    # conditioning vs the oracle on the worst pixels
    bad = torch.nonzero(e.view(-1) > 0.01).view(-1)[:64].cpu()
    print("bad pixels (first)", bad[:16].tolist())
    if bad.numel():
        c32, _ = ctx.gather_cossim(sc, S, ray_idx=bad)
        fl = [rf[0][0].permute(0, 2, 3, 1).contiguous().cpu().half().float(), rf[1][0].permute(0, 2, 3, 1).contiguous().cpu().half().float()]
        ex, it, nf = cpu(var.extrinsics), cpu(var.intrinsics), cpu(var.near_fars)
        aux = RO.render_rays(dec_sd, fl, imgs[0].permute(0, 2, 3, 1).contiguous().cpu(), ex[0, :3, :3], it[0, :3], nf[0, :3], ex[0, 3, :3], it[0, 3], nf[0, 3], bad, S, return_aux=True)
        cond = aux[3]["cond"]
        d = (c32.cpu() - cond).abs()
        print("cond max err per column", [round(float(x), 4) for x in d.amax(0)])
        print("oracle rgb vs reference rgb on bad pixels", rmsf(aux[0], r.rgb[0][bad].cpu()), " ours", rmsf(o[0][bad].cpu(), r.rgb[0][bad].cpu()))
        ndc = aux[3]["ndc"]; print("ndc range", ndc.amin(0).tolist(), ndc.amax(0).tolist())
