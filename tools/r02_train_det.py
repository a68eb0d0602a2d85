"""Run-to-run variation of the training gradients (same model, same seed, same batch): per-part relative difference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matchnerf_b200.matchnerf import MatchNeRF
from matchnerf_b200.utils import AttrDict
from oracle import synth
from tests.test_host_cpu import make_opts

DEV = "cuda:0"
H, W, S, R = 64, 96, 16, 256
opt = make_opts(**{"nerf.sample_intvs": S, "nerf.rand_rays_train": R, "nerf.sample_stratified": False})
opt.device = DEV
m = MatchNeRF(opt).train()
m.feat_enc.load_state_dict(synth.synthetic_encoder(1)); m.nerf_dec.load_state_dict(synth.synthetic_decoder(0)); m.to(DEV)
g = torch.Generator().manual_seed(21)
images = torch.rand(1, 4, 3, H, W, generator=g)
extr, intr, nf = synth.synthetic_cameras(H, W)
batch = lambda: AttrDict(images=images.to(DEV), extrinsics=extr.to(DEV), intrinsics=intr.to(DEV), near_fars=nf.to(DEV))
runs = []
for r in range(3):
    torch.manual_seed(100)
    m.zero_grad(set_to_none=True)
    out = m(batch(), mode="train")
    gt = images[0, 3].permute(1, 2, 0).reshape(-1, 3).to(DEV)[out["ray_idx"]]
    loss = torch.nn.functional.mse_loss(out["rgb"][0], gt)
    loss.backward()
    runs.append((float(loss), out["ray_idx"].clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
print("losses", [r[0] for r in runs], "same rays", torch.equal(runs[0][1], runs[1][1]))
for part in ("feat_enc.backbone", "feat_enc.transformer", "feat_enc.featup", "nerf_dec"):
    ks = [k for k in runs[0][2] if k.startswith(part)]
    a = torch.cat([runs[0][2][k].reshape(-1) for k in ks]); b = torch.cat([runs[1][2][k].reshape(-1) for k in ks]); c = torch.cat([runs[2][2][k].reshape(-1) for k in ks])
    print(part, "norm %.4e" % float(a.norm()), "rel diff run0-1 %.3e" % float((a - b).norm() / a.norm()), "run1-2 %.3e" % float((b - c).norm() / b.norm()))
