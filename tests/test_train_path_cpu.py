"""CPU: the torch-side half of the training step (matchnerf_b200/train_path.py: decoder on flat sample lists, compositing, the
index-gathered split-window attention) against the oracle -- values AND gradients, since these functions exist to be differentiated."""
import torch

from matchnerf_b200 import train_path as TP
from matchnerf_b200.cond_nerf import CondNeRF
from oracle import encoder_oracle as EO
from oracle import render_oracle as RO
from oracle import synth
from tests.test_host_cpu import make_opts


def _decoder_inputs(R, S, seed=0):
    g = torch.Generator().manual_seed(seed)
    ndc = torch.randn(R, S, 3, generator=g) * 0.5
    dirs = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    cond = torch.randn(R * S, 22, generator=g) * 0.3
    cond[:, 19:22] = (torch.rand(R * S, 3, generator=g) > 0.4).float()
    return ndc, dirs, cond


def test_decode_and_composite_match_oracle_with_gradients():
    for act, posenc, maskfill in (("ReLU", False, False), ("ELU", True, True)):
        opt = make_opts(**{"nerf.sample_intvs": 16, "decoder.raytrans_act": act, "decoder.raytrans_posenc": posenc,
                           "decoder.density_maskfill": maskfill})
        dec = CondNeRF(opt)
        sd = synth.synthetic_decoder(0)
        dec.load_state_dict(sd)
        R, S = 6, 16
        ndc, dirs, cond = _decoder_inputs(R, S)
        cond_a = cond.clone().requires_grad_(True)
        rgb, sig = TP.decode_samples(dec, opt, ndc, dirs, cond_a, S)
        depth = torch.rand(R, S).sort(-1).values
        out = TP.composite_samples(rgb, sig, depth, True)
        sd_g = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        cond_b = cond.clone().requires_grad_(True)
        o_rgb, o_sig = RO.decoder(sd_g, ndc.reshape(R * S, 3), dirs, cond_b, S, raytrans_act=act, raytrans_posenc=posenc,
                                  density_maskfill=maskfill)
        ref = RO.composite(o_sig.reshape(R, S), o_rgb.reshape(R, S, 3), depth, True)
        assert float(sig.detach().mean()) > 1e-3
        for a, b in zip(out, ref[:3]):
            assert float((a - b.reshape(a.shape)).abs().max()) < 1e-5
        tgt = torch.rand(R, 3, generator=torch.Generator().manual_seed(5))
        ((out[0] - tgt) ** 2).mean().backward()
        ((ref[0] - tgt) ** 2).mean().backward()
        assert float((cond_a.grad - cond_b.grad).abs().max()) < 1e-6 * max(1.0, float(cond_b.grad.abs().max()))
        for k, p in dec.state_dict(keep_vars=True).items():
            assert float((p.grad - sd_g[k].grad).abs().max()) <= 1e-5 * max(1e-3, float(sd_g[k].grad.abs().max())), k


def test_window_attention_autograd_matches_oracle():
    g = torch.Generator().manual_seed(1)
    for (h, w, k, shift) in ((8, 12, 2, False), (8, 12, 2, True), (12, 16, 4, True), (6, 10, 1, False)):
        q, kk, v = (torch.randn(2, h * w, 128, generator=g).requires_grad_(True) for _ in range(3))
        a = TP.window_attention_autograd(q, kk, v, h, w, k, shift)
        b = EO.window_attention(q, kk, v, h, w, k, shift)
        assert float((a - b).abs().max()) < 1e-5
        wgt = torch.randn(a.shape, generator=g)
        ga = torch.autograd.grad((a * wgt).sum(), (q, kk, v))
        gb = torch.autograd.grad((b * wgt).sum(), (q, kk, v))
        for x, y in zip(ga, gb):
            assert float((x - y).abs().max()) < 1e-5


def test_sample_geometry_matches_oracle():
    """ndc / view direction / depth of the training path vs the oracle's ray casting + projection (float64-inverted pose)."""
    from matchnerf_b200 import capi
    H, W, S = 40, 56, 8
    extr, intr, nf = synth.synthetic_cameras(H, W)
    z = torch.zeros(1)                                     # (no feature maps needed: only the camera block is read)
    sc = capi.PackedScene(z, z, z, H, W, 5, 7, 10, 14, extr[0, :3, :3].contiguous(), intr[0, :3].contiguous(),
                          nf[0, :3].contiguous()).c_scene(extr[0, 3, :3], intr[0, 3], nf[0, 3])
    idx = torch.tensor([0, 5, W * 7 + 3, H * W - 1])
    jit = torch.rand(idx.numel(), S, generator=torch.Generator().manual_seed(2))
    ndc, dirs, depth = TP.sample_geometry(sc, idx, jit, S, "cpu")
    centre, ray = RO.cast_rays(H, W, extr[0, 3, :3], intr[0, 3], idx)
    t = RO.sample_depths(float(nf[0, 3, 0]), float(nf[0, 3, 1]), S, jit)
    pts = (centre[None, None] + ray[:, None] * t[..., None]).reshape(-1, 3)
    ref = RO.project_ndc(pts, extr[0, 0, :3], intr[0, 0], W, H, float(nf[0, 0, 0]), float(nf[0, 0, 1]))
    assert float((ndc.reshape(-1, 3) - ref).abs().max()) < 2e-5
    assert float((depth - t).abs().max()) < 1e-6
    unit = ray / ray.norm(dim=-1, keepdim=True)
    assert float((dirs - unit @ extr[0, 0, :3, :3].T).abs().max()) < 1e-5
