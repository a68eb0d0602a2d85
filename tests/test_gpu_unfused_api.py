"""GPU (B200): the reference's unfused public per-sample methods -- MatchNeRF.query_cond_info (models/matchnerf.py:209),
CondNeRF.forward (models/rfdecoder/cond_nerf.py:52) and NeRF.composite (models/rfdecoder/nerf.py:101) -- called with
the reference's argument shapes, against the CPU oracle.  Chained together they must reproduce MatchNeRF.render."""
import pytest
import torch

from matchnerf_b200.utils import AttrDict
from oracle import render_oracle as RO
from oracle import synth
from tests.helpers import half_round, oracle_render, rms
from tests.test_host_cpu import make_opts

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(S, **over):
    from matchnerf_b200.matchnerf import MatchNeRF
    opt = make_opts(**{"nerf.sample_intvs": S, **over})
    opt.device = DEV
    m = MatchNeRF(opt).eval()
    m.nerf_dec.load_state_dict(synth.synthetic_decoder(0), strict=True)
    return m.to(DEV), opt


def scene(H=64, W=96, seed=11, baseline=10.0):
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(1, 3, 256, H // 8, W // 8, generator=g), torch.randn(1, 3, 256, H // 4, W // 4, generator=g)]
    imgs = torch.rand(1, 3, 3, H, W, generator=g)
    extr, intr, nf = synth.synthetic_cameras(H, W, baseline_deg=baseline)
    ref = dict(extrinsics=extr[:, :3, :3].to(DEV), intrinsics=intr[:, :3].to(DEV), near_fars=nf[:, :3].to(DEV))
    tgt = dict(extrinsics=extr[:, 3, :3].to(DEV), intrinsics=intr[:, 3].to(DEV), near_fars=nf[:, 3].to(DEV))
    return feats, imgs, extr, intr, nf, ref, tgt


@pytest.mark.parametrize("S,R,baseline", [(16, 50, 10.0), (5, 33, 25.0), (64, 4, 10.0)])
def test_unfused_chain_matches_oracle_and_fused_render(S, R, baseline):
    m, opt = build(S)
    feats, imgs, extr, intr, nf, ref, tgt = scene(baseline=baseline)
    H, W = imgs.shape[-2:]
    ray_idx = torch.randperm(H * W, generator=torch.Generator().manual_seed(S))[:R]
    o_rgb, o_depth, o_op, aux = oracle_render(synth.synthetic_decoder(0), feats, imgs, extr, intr, nf, ray_idx, S,
                                              quantize_feats=True, return_aux=True)
    fd = [f.to(DEV) for f in feats]
    im = imgs.to(DEV)
    with torch.no_grad():
        # 1. query_cond_info on the oracle's world-space sample points, reference shapes [B,R,S,3]
        pts = aux["pts"].reshape(1, R, S, 3).to(DEV)
        cond = m.query_cond_info(pts, ref, im, fd)
        assert cond["feat_info"].shape == (1, R, S, 10) and cond["color_info"].shape == (1, R, S, 9) and cond["mask_info"].shape == (1, R, S, 3)
        got = torch.cat([cond["feat_info"], cond["color_info"], cond["mask_info"]], -1).reshape(R * S, 22)
        assert rms(got[:, 10:19], aux["cond"][:, 10:19]) < 2e-5
        assert rms(got[:, :10], aux["cond"][:, :10]) < 5e-4
        assert float((got[:, 19:].cpu() != aux["cond"][:, 19:]).float().mean()) < 5e-3
        # 2. CondNeRF.forward on explicit tensors (oracle conditioning, so the comparison isolates the decoder)
        ocond = aux["cond"].reshape(1, R, S, 22).to(DEV)
        cinfo = {"feat_info": ocond[..., :10], "color_info": ocond[..., 10:19], "mask_info": ocond[..., 19:]}
        ndc = aux["ndc"].reshape(1, R, S, 3).to(DEV)
        dirs = aux["dir_ref"].to(DEV)[None, :, None, :].expand(1, R, S, 3)
        rgb_s, sigma = m.nerf_dec(opt, ndc, ray_unit=dirs, cond_info=cinfo)
        assert rgb_s.shape == (1, R, S, 3) and sigma.shape == (1, R, S)
        assert rms(rgb_s.reshape(-1, 3), aux["rgb_s"]) < 2e-5 and rms(sigma.reshape(-1), aux["sigma"]) < 1e-4 * max(1.0, float(aux["sigma"].abs().max()))
        # 3. composite on explicit tensors, both background modes
        depth_s = aux["depth_samples"].reshape(1, R, S, 1).to(DEV)
        for bg in (False, True):
            rgb, depth, op, prob = m.nerf_dec.composite(opt, None, rgb_s, sigma, depth_s, bg)
            e = RO.composite(aux["sigma"].reshape(R, S), aux["rgb_s"].reshape(R, S, 3), aux["depth_samples"], bg)
            assert rgb.shape == (1, R, 3) and depth.shape == (1, R, 1) and op.shape == (1, R, 1) and prob.shape == (1, R, S, 1)
            assert rms(rgb[0], e[0]) < 2e-5 and rms(depth[0], e[1]) < 5e-5 and rms(op[0], e[2]) < 2e-5 and rms(prob[0, ..., 0], e[3]) < 2e-5
        # 4. the chain equals the fused render on the same rays (fp32 decoder both ways -> fp32 round-off only)
        fused = m.render(opt, tgt, ray_idx=ray_idx.to(DEV), mode="test", ref_poses=ref, ref_images=im, ref_feats_list=fd)
        rgb, depth, op, _ = m.nerf_dec.composite(opt, None, rgb_s, sigma, depth_s, False)
        assert rms(fused.rgb, rgb) < 2e-3 and rms(fused.opacity, op) < 4e-3      # fused path runs the fp16-operand tcgen05 decoder
    assert 0.02 < float(o_op.mean()) < 0.98


def test_composite_long_ray_and_empty():
    from matchnerf_b200 import capi
    ctx = capi.get_context(torch.device(DEV))
    g = torch.Generator().manual_seed(2)
    R, S = 37, 200                                        # > 32 samples: the scan carries across warp steps
    sigma = torch.rand(R, S, generator=g) * 0.1
    rgb = torch.rand(R, S, 3, generator=g)
    depth = torch.rand(R, S, generator=g) * 3 + 2
    got = ctx.composite(rgb.to(DEV), sigma.to(DEV), depth.to(DEV), True)
    e = RO.composite(sigma, rgb, depth, True)
    assert rms(got[0], e[0]) < 2e-5 and rms(got[1][:, None], e[1]) < 5e-5 and rms(got[2][:, None], e[2]) < 2e-5 and rms(got[3], e[3]) < 2e-5
    z = ctx.composite(torch.zeros(0, S, 3, device=DEV), torch.zeros(0, S, device=DEV), torch.zeros(0, S, device=DEV))
    assert z[0].shape == (0, 3)


def test_unfused_methods_fail_loudly_on_cpu_tensors():
    m, opt = build(8)
    with pytest.raises(RuntimeError):
        m.nerf_dec.composite(opt, None, torch.zeros(1, 2, 8, 3), torch.zeros(1, 2, 8), torch.zeros(1, 2, 8, 1), False)
    with pytest.raises(RuntimeError):
        m.query_cond_info(torch.zeros(1, 2, 8, 3), None, None, None)
