"""GPU (B200): parity of each CUDA kernel, called through the C ABI, against the CPU oracle on seeded inputs and
against the golden vectors of the unmodified reference.

Tolerances (fp32 reference arithmetic; the kernels read fp16 feature maps):
  * kernel vs oracle evaluated on the SAME fp16-rounded feature maps: fp32 round-off only (<= 2e-5 RMS) for colours,
    masks and the fp32 decoder; <= 5e-4 RMS for the cosine similarities (the 4 taps are blended in packed fp16);
  * kernel vs reference goldens (fp32 feature maps): the north-star budget, rgb RMS <= 2e-3 (== 0.01 dB at 27 dB
    PSNR) -- measured values are ~1e-4.
"""
import os

import numpy as np
import pytest
import torch

from oracle import encoder_oracle as EO
from oracle import render_oracle as RO
from oracle import synth
from tests.helpers import (config1_inputs, dec_from_npz, frac_above, half_round, load_npz, max_abs, oracle_render, psnr, rms)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GATHER_IMPL_DEFAULT = 3       # csrc/gather.cu gather_impl()


def make_scene(ctx, feats, imgs, extr, intr, nf, local_radius=0, local_dilation=1):
    packed = ctx.pack_scene([feats[0][0].to(DEV), feats[1][0].to(DEV)], imgs[0].to(DEV), extr[0, :3], intr[0, :3], nf[0, :3],
                            local_radius, local_dilation)
    return packed, packed.c_scene(extr[0, 3, :3], intr[0, 3], nf[0, 3])


def make_cfg(S, act="ReLU", posenc=False, maskfill=False):
    from matchnerf_b200 import capi
    cfg = capi.DecoderCfg()
    cfg.n_samples, cfg.raytrans_act, cfg.raytrans_posenc, cfg.density_maskfill = S, {"ReLU": 0, "ELU": 1}[act], int(posenc), int(maskfill)
    return cfg


# ------------------------------------------------------------------------------------------- tcgen05 building blocks
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N,K", [(128, 64), (128, 128), (128, 192), (64, 128), (16, 64), (80, 128), (256, 64)])
def test_umma_selftest(ctx, mode, N, K):
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    a = torch.randn(128, K, generator=g).half()
    b = torch.randn(N, K, generator=g).half()
    d = ctx.selftest_umma(a, b, mode).cpu()
    ref = (a.double() @ b.double().T).float()
    assert max_abs(d, ref) < 2e-3 * max(1.0, float(ref.abs().max())), (mode, N, K, max_abs(d, ref))


@pytest.mark.parametrize("N,K", [(128, 128), (64, 64), (128, 64), (256, 128)])
def test_umma_selftest_mn_major_b(ctx, N, K):
    """B consumed as an MN-major operand (the V tile of attention: rows = keys, contiguous channels)."""
    g = torch.Generator().manual_seed(N + K)
    a = torch.randn(128, K, generator=g).half()
    b = torch.randn(K, N, generator=g).half()
    d = ctx.selftest_umma(a, b, 2).cpu()
    ref = (a.double() @ b.double()).float()
    assert max_abs(d, ref) < 2e-3 * max(1.0, float(ref.abs().max())), max_abs(d, ref)


# ------------------------------------------------------------------------------------------- K-gather
def test_pack_features_layout(ctx):
    """The packed fp16 layout the gather kernel reads (DESIGN.md "feature map layout"), for an odd and an even width."""
    g = torch.Generator().manual_seed(3)
    f = torch.randn(3, 256, 5, 7, generator=g)
    f1 = torch.randn(3, 256, 10, 14, generator=g)
    imgs = torch.rand(1, 3, 3, 40, 56, generator=g)
    packed, _ = make_scene(ctx, [f[None], f1[None]], imgs, *synth.synthetic_cameras(40, 56))
    impl = os.environ.get("MNF_GATHER_IMPL", str(GATHER_IMPL_DEFAULT))
    for src, buf in ((f, packed.feat0), (f1, packed.feat1)):
        V, _, h, w = src.shape
        want = src.half().float().permute(0, 2, 3, 1)                      # [V,h,w,256] natural channel order
        flat = buf.cpu().float()
        if impl == "4":                                                    # x-pair interleaved blocks [V][h][ceil(w/2)][256][2]
            wp = (w + 1) // 2
            body = flat[: V * h * wp * 512].view(V, h, wp, 256, 2)
            got = body.permute(0, 1, 2, 4, 3).reshape(V, h, 2 * wp, 256)
            assert torch.equal(got[:, :, :w], want)
            assert float(got[:, :, w:].abs().max() if 2 * wp > w else 0.0) == 0.0      # unpaired texel of an odd width
            assert float(flat[V * h * wp * 512: V * h * wp * 512 + (wp + 4) * 512].abs().max()) == 0.0   # zero tail
        else:
            p = flat[: V * h * w * 256].view(V, h, w, 256)
            pos = torch.arange(256)                                        # v3 packing: slot l = channels 4l..4l+3 of both halves
            lane, e = pos // 8, pos % 8
            chan = torch.where(e < 4, 0, 128) + 4 * lane + (e & 3)
            assert torch.equal(p, want[..., chan])
    assert torch.equal(packed.images.cpu()[..., :3].permute(0, 3, 1, 2), imgs[0])


SMALL = ["small_base", "small_wide_baseline", "small_elu_maskfill_posenc_bg", "small_s24_elu_maskfill_posenc_bg", "demo_own_S128",
         "video_own_S256"]


@pytest.mark.parametrize("name", SMALL)
def test_gather_small_golden(ctx, golden_dir, name):
    z = load_npz(golden_dir, name + ".npz")
    S = int(z["S"])
    feats = [torch.from_numpy(z["feat8"]), torch.from_numpy(z["feat4"])]
    imgs, extr, intr, nf = (torch.from_numpy(z[k]) for k in ("images", "extrinsics", "intrinsics", "near_fars"))
    ray_idx = torch.from_numpy(z["ray_idx"])
    packed, sc = make_scene(ctx, feats, imgs, extr, intr, nf)
    c32, c16 = ctx.gather_cossim(sc, S, ray_idx=ray_idx, want_f32=True, want_f16=True)
    torch.cuda.synchronize()
    aux = oracle_render(dec_from_npz(z), feats, imgs, extr, intr, nf, ray_idx, S, quantize_feats=True, return_aux=True)[3]
    cond_q = aux["cond"]
    # masks and colours do not depend on the feature quantisation; sims compared against the quantised-feature oracle
    assert frac_above(c32[:, 19:], cond_q[:, 19:], 0.5) < 2e-3          # a mask flips only for samples on the border
    assert rms(c32[:, 10:19], cond_q[:, 10:19]) < 2e-5
    assert rms(c32[:, :10], cond_q[:, :10]) < 5e-4, rms(c32[:, :10], cond_q[:, :10])   # taps blended in packed fp16
    assert rms(c32[:, :10], z["cond"][:, :10]) < 1e-3                   # vs the reference on fp32 features
    assert rms(c16[:, :22].float(), c32) < 5e-4 and float(c16[:, 22:].abs().max()) == 0.0


@pytest.mark.parametrize("radius,dilation", [(1, 2), (2, 1)])
def test_gather_local_radius(ctx, golden_dir, radius, dilation):
    """encoder.feature_sample_local_radius > 0 (models/gmflow/utils.py:136-162): mean over (2r+1)^2 dilated bilinear samples, with
    the reference's re-normalisation.  (1, 2) = the golden of the unmodified reference; every ray-input kind (explicit ids,
    contiguous range, explicit points) goes through the same kernel; the fused render agrees with the reference's rgb."""
    z = load_npz(golden_dir, "small_local_radius1.npz")
    S = int(z["S"])
    feats = [torch.from_numpy(z["feat8"]), torch.from_numpy(z["feat4"])]
    imgs, extr, intr, nf = (torch.from_numpy(z[k]) for k in ("images", "extrinsics", "intrinsics", "near_fars"))
    ray_idx = torch.from_numpy(z["ray_idx"])
    packed, sc = make_scene(ctx, feats, imgs, extr, intr, nf, radius, dilation)
    c32, c16 = ctx.gather_cossim(sc, S, ray_idx=ray_idx, want_f32=True, want_f16=True)
    torch.cuda.synchronize()
    dec = dec_from_npz(z)
    o = oracle_render(dec, feats, imgs, extr, intr, nf, ray_idx, S, quantize_feats=True, return_aux=True,
                      local_radius=radius, local_dilation=dilation)
    cond_q = o[3]["cond"]
    assert frac_above(c32[:, 19:], cond_q[:, 19:], 0.5) < 2e-3 and rms(c32[:, 10:19], cond_q[:, 10:19]) < 2e-5
    assert rms(c32[:, :10], cond_q[:, :10]) < 2e-5, rms(c32[:, :10], cond_q[:, :10])      # fp32 blend: round-off only
    assert rms(c16[:, :22].float(), c32) < 5e-4 and float(c16[:, 22:].abs().max()) == 0.0
    plain = oracle_render(dec, feats, imgs, extr, intr, nf, ray_idx, S, quantize_feats=True, return_aux=True)[3]["cond"]
    assert rms(cond_q[:, :10], plain[:, :10]) > 1e-2                                    # the option changes the answer
    if (radius, dilation) == (int(z["local_radius"]), int(z["local_dilation"])):
        assert rms(c32[:, :10], z["cond"][:, :10]) < 1e-3                               # vs the reference on fp32 features
        ctx.load_decoder(dec)
        rgb, depth, op = ctx.render_rays(sc, make_cfg(S), ray_idx=ray_idx, impl=2)
        assert rms(rgb, z["rgb"]) < 2e-3 and rms(op, z["opacity"][:, 0]) < 2e-3
    # contiguous range and explicit points: the same kernel, identical bits
    first = 40 * 7 + 3
    a, _ = ctx.gather_cossim(sc, S, first_ray=first, n_rays=50)
    b, _ = ctx.gather_cossim(sc, S, ray_idx=torch.arange(first, first + 50))
    assert torch.equal(a, b)
    pts = o[3]["pts"].reshape(-1, S, 3)[:64].contiguous()
    cpts, _ = ctx.query_cond_points(sc, pts.to(DEV))
    assert torch.equal(cpts, c32[: 64 * S])


def test_gather_config1_and_ray_range(ctx):
    feats, imgs, extr, intr, nf, ray_idx = config1_inputs()
    packed, sc = make_scene(ctx, feats, imgs, extr, intr, nf)
    S = 64
    sub = ray_idx[:128]
    c32, _ = ctx.gather_cossim(sc, S, ray_idx=sub)
    aux = oracle_render(synth.synthetic_decoder(0), feats, imgs, extr, intr, nf, sub, S, quantize_feats=True, return_aux=True)[3]
    assert rms(c32[:, 10:], aux["cond"][:, 10:]) < 2e-5 and rms(c32[:, :10], aux["cond"][:, :10]) < 5e-4
    # contiguous range (tensor-core gather) vs explicit ids (v3 kernel): identical colours / masks, cosines to the blend precision
    first = 640 * 100 + 17
    a, _ = ctx.gather_cossim(sc, S, first_ray=first, n_rays=96)
    b, _ = ctx.gather_cossim(sc, S, ray_idx=torch.arange(first, first + 96))
    assert torch.equal(a[:, 10:], b[:, 10:]) and rms(a[:, :10], b[:, :10]) < 5e-4
    # the same range twice, and as part of a longer range: bit-identical (results do not depend on the slicing)
    a2, _ = ctx.gather_cossim(sc, S, first_ray=first, n_rays=96)
    a3, _ = ctx.gather_cossim(sc, S, first_ray=first - 40, n_rays=300)
    assert torch.equal(a, a2) and torch.equal(a, a3[40 * S: (40 + 96) * S])


@pytest.mark.parametrize("S,first,n", [(13, 0, 7), (64, 40 * 56 - 9, 9), (8, 56 * 17 + 50, 23), (2, 5, 1)])
def test_gather_borders_ragged(ctx, S, first, n):
    """Image corners (taps on the last row / column of the maps, clipped coordinates), sample counts that are not a
    multiple of the 8-sample chunk, ray counts that are not a multiple of the 4-ray quad; wide baseline so that many
    samples leave the source views."""
    H, W = 40, 56
    g = torch.Generator().manual_seed(S * 7 + n)
    feats = [torch.randn(1, 3, 256, H // 8, W // 8, generator=g), torch.randn(1, 3, 256, H // 4, W // 4, generator=g)]
    imgs = torch.rand(1, 3, 3, H, W, generator=g)
    extr, intr, nf = synth.synthetic_cameras(H, W, baseline_deg=25.0)
    packed, sc = make_scene(ctx, feats, imgs, extr, intr, nf)
    c32, c16 = ctx.gather_cossim(sc, S, first_ray=first, n_rays=n, want_f32=True, want_f16=True)
    torch.cuda.synchronize()
    ray_idx = torch.arange(first, first + n)
    aux = oracle_render(synth.synthetic_decoder(0), feats, imgs, extr, intr, nf, ray_idx, S, quantize_feats=True, return_aux=True)[3]
    cond = aux["cond"]
    assert c32.shape == (n * S, 22) and c16.shape == (n * S, 32)
    assert 0.02 < float(cond[:, 19:].mean()) < 0.98                   # both inside and outside samples are present
    assert frac_above(c32[:, 19:], cond[:, 19:], 0.5) < 5e-3
    assert rms(c32[:, 10:19], cond[:, 10:19]) < 2e-5
    assert rms(c32[:, :10], cond[:, :10]) < 5e-4, rms(c32[:, :10], cond[:, :10])
    assert rms(c16[:, :22].float(), c32) < 5e-4 and float(c16[:, 22:].abs().max()) == 0.0
    b32, _ = ctx.gather_cossim(sc, S, ray_idx=ray_idx, want_f32=True)          # explicit ray list: the v3 kernel
    assert torch.equal(b32[:, 10:], c32[:, 10:]) and rms(b32[:, :10], c32[:, :10]) < 5e-4
    assert rms(b32[:, :10], cond[:, :10]) < 5e-4


@pytest.mark.parametrize("H,W,S,first,n,baseline", [
    (512, 640, 64, 640 * 96, 640 * 40, 10.0),       # whole 8-row bands of the DTU image
    (512, 640, 16, 640 * 101 + 37, 640 * 9 + 5, 10.0),   # ragged start / end inside rows, partial bands
    (40, 56, 13, 0, 40 * 56, 10.0),                  # width not a multiple of the 16-pixel tile, tiny feature maps
    (64, 96, 8, 96 * 3, 96 * 50, 30.0),              # wide baseline: footprints leave the boxes -> the v3 fix-up pass runs
    (100, 72, 5, 11, 100 * 72 - 30, 10.0),           # height not a multiple of the 8-row band
])
def test_gather_tensor_core_path(ctx, H, W, S, first, n, baseline):
    """Contiguous ray ranges of >= 1024 rays take the tcgen05 + TMA gather (csrc/gather_tc.cu, with the v3 kernel as its fix-up
    pass); the same rays given as an explicit list take the v3 kernel.  Both against the CPU oracle: colours and masks to fp32
    round-off, cosine similarities to the fp16-feature tolerance; and against each other."""
    g = torch.Generator().manual_seed(H * 7 + S)
    feats = [torch.randn(1, 3, 256, H // 8, W // 8, generator=g), torch.randn(1, 3, 256, H // 4, W // 4, generator=g)]
    imgs = torch.rand(1, 3, 3, H, W, generator=g)
    extr, intr, nf = synth.synthetic_cameras(H, W, baseline_deg=baseline)
    packed, sc = make_scene(ctx, feats, imgs, extr, intr, nf)
    t32, t16 = ctx.gather_cossim(sc, S, first_ray=first, n_rays=n, want_f32=True, want_f16=True)
    v32, _ = ctx.gather_cossim(sc, S, ray_idx=torch.arange(first, first + n), want_f32=True)
    torch.cuda.synchronize()
    assert t32.shape == (n * S, 22) and t16.shape == (n * S, 32)
    assert torch.equal(t32[:, 10:], v32[:, 10:])                       # colours and masks: the same arithmetic in both kernels
    assert rms(t32[:, :10], v32[:, :10]) < 5e-4
    assert rms(t16[:, :22].float(), t32) < 5e-4 and float(t16[:, 22:].abs().max()) == 0.0
    sub = torch.arange(first, first + n, max(1, n // 300))              # the oracle on ~300 rays spread over the range
    aux = oracle_render(synth.synthetic_decoder(0), feats, imgs, extr, intr, nf, sub, S, quantize_feats=True, return_aux=True)[3]
    rows = ((sub - first)[:, None] * S + torch.arange(S)[None]).reshape(-1)
    got = t32[rows.to(DEV)]
    assert frac_above(got[:, 19:], aux["cond"][:, 19:], 0.5) < 5e-3
    assert rms(got[:, 10:19], aux["cond"][:, 10:19]) < 2e-5
    assert rms(got[:, :10], aux["cond"][:, :10]) < 5e-4, rms(got[:, :10], aux["cond"][:, :10])


# ------------------------------------------------------------------------------------------- K-mlp-composite
def run_decoder_case(ctx, dec, feats, imgs, extr, intr, nf, ray_idx, S, impl, act="ReLU", posenc=False, maskfill=False, bg=False):
    ctx.load_decoder(dec)
    packed, sc = make_scene(ctx, feats, imgs, extr, intr, nf)
    cfg = make_cfg(S, act, posenc, maskfill)
    o = oracle_render(dec, feats, imgs, extr, intr, nf, ray_idx, S, quantize_feats=True, return_aux=True, setbg_opaque=bg,
                      raytrans_act=act, raytrans_posenc=posenc, density_maskfill=maskfill)
    cond = o[3]["cond"].to(DEV)
    cond16 = torch.zeros(cond.shape[0], 32, dtype=torch.float16, device=DEV)
    cond16[:, :22] = cond.half()
    rgb, depth, op, aux = ctx.decoder_composite(sc, cfg, cond_f32=cond, cond_f16=cond16, ray_idx=ray_idx, setbg_opaque=bg,
                                                impl=impl, want_aux=True)
    torch.cuda.synchronize()
    return (rgb, depth, op, aux), o


@pytest.mark.parametrize("name", SMALL)
def test_decoder_fp32_small_golden(ctx, golden_dir, name):
    z = load_npz(golden_dir, name + ".npz")
    feats = [torch.from_numpy(z["feat8"]), torch.from_numpy(z["feat4"])]
    imgs, extr, intr, nf = (torch.from_numpy(z[k]) for k in ("images", "extrinsics", "intrinsics", "near_fars"))
    got, o = run_decoder_case(ctx, dec_from_npz(z), feats, imgs, extr, intr, nf, torch.from_numpy(z["ray_idx"]), int(z["S"]), 1,
                              act=str(z["raytrans_act"]), posenc=bool(z["raytrans_posenc"]), maskfill=bool(z["density_maskfill"]),
                              bg=bool(z["setbg_opaque"]))
    assert rms(got[3][:, :3], o[3]["rgb_s"]) < 2e-5 and rms(got[3][:, 3], o[3]["sigma"]) < 2e-5
    assert rms(got[0], o[0]) < 2e-5 and rms(got[1], o[1][:, 0]) < 1e-4 and rms(got[2], o[2][:, 0]) < 2e-5


@pytest.mark.parametrize("S", [64, 128])
def test_decoder_fp32_config1(ctx, S):
    feats, imgs, extr, intr, nf, ray_idx = config1_inputs()
    got, o = run_decoder_case(ctx, synth.synthetic_decoder(0), feats, imgs, extr, intr, nf, ray_idx[:256], S, 1)
    assert rms(got[0], o[0]) < 2e-5 and rms(got[1], o[1][:, 0]) < 1e-4 and rms(got[2], o[2][:, 0]) < 2e-5


TC_SAMPLES = (16, 32, 64, 128, 256)      # sample counts the tcgen05 decoder covers (csrc/decoder_tc.cu decoder_tc_supports)


@pytest.mark.parametrize("S", TC_SAMPLES)
def test_decoder_tcgen05_config1(ctx, S):
    """fp16-operand tensor-core kernel vs the fp32 oracle: mixed-precision budget rgb RMS <= 2e-3 (0.01 dB).
    A sample count in TC_SAMPLES that the library rejects is a FAILURE (no skip: the target box must run it)."""
    feats, imgs, extr, intr, nf, ray_idx = config1_inputs()
    got, o = run_decoder_case(ctx, synth.synthetic_decoder(0, density_gain=(1.0 if S <= 128 else 0.25)), feats, imgs, extr, intr, nf,
                              ray_idx[:512], S, 2)
    e_rgb, e_depth, e_op = rms(got[0], o[0]), rms(got[1], o[1][:, 0]), rms(got[2], o[2][:, 0])
    assert e_rgb < 1e-3 and e_depth < 6e-3 and e_op < 2e-3, (e_rgb, e_depth, e_op)


@pytest.mark.parametrize("name", ["small_base", "small_wide_baseline", "small_elu_maskfill_posenc_bg", "small_s24_elu_maskfill_posenc_bg",
                                  "demo_own_S128", "video_own_S256"])
def test_decoder_tcgen05_small_golden(ctx, golden_dir, name):
    """Every option of the shipped configs on the TENSOR-CORE kernel: ELU, raytrans_posenc, density_maskfill, white background
    (configs/demo_own.yaml:10-15 at S = 128, configs/test_video_own.yaml:10-15 at S = 256, train_ibrnet.yaml:12-14).
    The S = 24 case is outside the kernel's tiling (S must divide 128 or be 256): there the library must REFUSE impl = 2."""
    z = load_npz(golden_dir, name + ".npz")
    feats = [torch.from_numpy(z["feat8"]), torch.from_numpy(z["feat4"])]
    imgs, extr, intr, nf = (torch.from_numpy(z[k]) for k in ("images", "extrinsics", "intrinsics", "near_fars"))
    args = (ctx, dec_from_npz(z), feats, imgs, extr, intr, nf, torch.from_numpy(z["ray_idx"]), int(z["S"]), 2)
    kw = dict(act=str(z["raytrans_act"]), posenc=bool(z["raytrans_posenc"]), maskfill=bool(z["density_maskfill"]), bg=bool(z["setbg_opaque"]))
    if int(z["S"]) not in TC_SAMPLES:
        with pytest.raises(RuntimeError, match="does not cover"):
            run_decoder_case(*args, **kw)
        return
    got, o = run_decoder_case(*args, **kw)
    assert rms(got[3][:, :3], o[3]["rgb_s"]) < 2e-3 and rms(got[3][:, 3], o[3]["sigma"]) < 2e-3 * max(1.0, float(o[3]["sigma"].abs().max()))
    assert rms(got[0], o[0]) < 1e-3 and rms(got[2], o[2][:, 0]) < 2e-3 and rms(got[1], o[1][:, 0]) < 6e-3
    if bool(z["density_maskfill"]):        # the masked samples are EXACTLY zero, as in cond_nerf.py:86-88
        unseen = (o[3]["cond"][:, 19:].sum(1) < 1).to(DEV)
        assert float(got[3][unseen, 3].abs().max() if bool(unseen.any()) else 0.0) == 0.0


@pytest.mark.parametrize("name", ["small_elu_maskfill_posenc_bg", "demo_own_S128", "video_own_S256"])
@pytest.mark.parametrize("impl", [1, 2])
def test_render_options_vs_reference_golden(ctx, golden_dir, name, impl):
    """Fused render (gather + decoder + composite through mnf_render_rays_fwd) with ELU / posenc / maskfill against the
    UNMODIFIED reference's outputs, on both decoder kernels."""
    z = load_npz(golden_dir, name + ".npz")
    feats = [torch.from_numpy(z["feat8"]), torch.from_numpy(z["feat4"])]
    imgs, extr, intr, nf = (torch.from_numpy(z[k]) for k in ("images", "extrinsics", "intrinsics", "near_fars"))
    ctx.load_decoder(dec_from_npz(z))
    packed, sc = make_scene(ctx, feats, imgs, extr, intr, nf)
    cfg = make_cfg(int(z["S"]), str(z["raytrans_act"]), bool(z["raytrans_posenc"]), bool(z["density_maskfill"]))
    rgb, depth, op = ctx.render_rays(sc, cfg, ray_idx=torch.from_numpy(z["ray_idx"]), setbg_opaque=bool(z["setbg_opaque"]), impl=impl)
    torch.cuda.synchronize()
    e = (rms(rgb, z["rgb"]), rms(depth, z["depth"][:, 0]), rms(op, z["opacity"][:, 0]))
    assert e[0] < 2e-3 and e[1] < 1.2e-2 and e[2] < 4e-3, e


# ------------------------------------------------------------------------------------------- fused render (C ABI) vs reference goldens
@pytest.mark.parametrize("S", [64, 128])
@pytest.mark.parametrize("impl", [1, 2])
def test_render_config1_vs_reference_golden(ctx, golden_dir, S, impl):
    """BASELINE config 1: 1024 rays x S samples x 3 views, outputs of the unmodified reference (fp32, CPU)."""
    z = load_npz(golden_dir, f"config1_synth_S{S}.npz")
    feats, imgs, extr, intr, nf, ray_idx = config1_inputs()
    ctx.load_decoder(synth.synthetic_decoder(0))
    packed, sc = make_scene(ctx, feats, imgs, extr, intr, nf)
    rgb, depth, op = ctx.render_rays(sc, make_cfg(S), ray_idx=ray_idx, impl=impl)
    torch.cuda.synchronize()
    e = (rms(rgb, z["rgb"]), rms(depth, z["depth"][:, 0]), rms(op, z["opacity"][:, 0]))
    assert e[0] < 2e-3 and e[1] < 1.2e-2 and e[2] < 4e-3, e
    # PSNR parity against a common pseudo ground truth (misc/metrics.py:35-41 formula): |delta| <= 0.01 dB
    gt = torch.from_numpy(z["rgb"]) + 0.045 * torch.randn(z["rgb"].shape, generator=torch.Generator().manual_seed(0))
    assert abs(psnr(rgb, gt) - psnr(z["rgb"], gt)) <= 0.01


def test_render_known_answer_reference_init(ctx, golden_dir):
    """SURVEY 8(c) known answer (reference default init): rgb mean 0.1042879 at S=64."""
    z = load_npz(golden_dir, "config1_refinit_S64.npz")
    feats, imgs, extr, intr, nf, ray_idx = config1_inputs()
    ctx.load_decoder(dec_from_npz(z))
    packed, sc = make_scene(ctx, feats, imgs, extr, intr, nf)
    rgb, depth, op = ctx.render_rays(sc, make_cfg(64), ray_idx=ray_idx, impl=1)
    assert abs(float(rgb.mean()) - 0.1042879) < 2e-4 and rms(rgb, z["rgb"]) < 1e-3


def test_render_edge_cases(ctx):
    feats, imgs, extr, intr, nf, ray_idx = config1_inputs()
    ctx.load_decoder(synth.synthetic_decoder(0))
    packed, sc = make_scene(ctx, feats, imgs, extr, intr, nf)
    # empty ray set
    rgb, depth, op = ctx.render_rays(sc, make_cfg(64), first_ray=0, n_rays=0, impl=1)
    assert rgb.shape == (0, 3)
    # ragged: a single ray, the last pixel; S not a multiple of anything
    rgb1, _, _ = ctx.render_rays(sc, make_cfg(37), first_ray=512 * 640 - 1, n_rays=1, impl=1)
    o = oracle_render(synth.synthetic_decoder(0), feats, imgs, extr, intr, nf, torch.tensor([512 * 640 - 1]), 37, quantize_feats=True)
    assert rms(rgb1, o[0]) < 3e-4
    # maximum S
    rgb2, _, op2 = ctx.render_rays(sc, make_cfg(256), ray_idx=ray_idx[:8], impl=1)
    o2 = oracle_render(synth.synthetic_decoder(0), feats, imgs, extr, intr, nf, ray_idx[:8], 256, quantize_feats=True)
    assert rms(rgb2, o2[0]) < 3e-4 and rms(op2, o2[2][:, 0]) < 5e-4
    # argument validation surfaces as errors, not crashes
    with pytest.raises(RuntimeError):
        ctx.render_rays(sc, make_cfg(300), ray_idx=ray_idx[:8])
    with pytest.raises(RuntimeError):
        ctx.render_rays(sc, make_cfg(64), first_ray=512 * 640 - 4, n_rays=8)
    # stratified sampling (train mode): explicit jitter
    jit = torch.rand(16, 64, generator=torch.Generator().manual_seed(2))
    rgb3, _, _ = ctx.render_rays(sc, make_cfg(64), ray_idx=ray_idx[:16], jitter=jit.to(DEV), impl=1)
    o3 = oracle_render(synth.synthetic_decoder(0), feats, imgs, extr, intr, nf, ray_idx[:16], 64, quantize_feats=True, jitter=jit)
    assert rms(rgb3, o3[0]) < 3e-4


# ------------------------------------------------------------------------------------------- K-attn
@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("name", ["window_attn_8x12_k2_s0", "window_attn_8x12_k2_s1", "window_attn_12x16_k4_s1",
                                  "window_attn_6x10_k1_s0"])
def test_window_attn_golden(ctx, golden_dir, name, impl):
    z = load_npz(golden_dir, name + ".npz")
    q, k, v = (torch.from_numpy(z[n]).to(DEV) for n in "qkv")
    out = ctx.window_attn(q, k, v, int(z["h"]), int(z["w"]), int(z["num_splits"]), bool(z["with_shift"]), impl=impl)
    tol = 2e-6 if impl == 1 else 2e-3
    assert rms(out, z["out"]) < tol, rms(out, z["out"])


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("shift", [False, True])
def test_window_attn_dtu_shape(ctx, impl, shift):
    """DTU shape: 64x80 tokens, 2x2 windows of L = 1280 (SURVEY 8a E-attn); oracle on one batch item."""
    g = torch.Generator().manual_seed(11)
    q, k, v = (torch.randn(2, 64 * 80, 128, generator=g) for _ in range(3))
    out = ctx.window_attn(q.to(DEV), k.to(DEV), v.to(DEV), 64, 80, 2, shift, impl=impl)
    ref = EO.window_attention(q[:1], k[:1], v[:1], 64, 80, 2, shift)
    tol = 2e-6 if impl == 1 else 2e-3
    assert rms(out[:1], ref) < tol
    # linearity in V (size-independent property): attn(q,k,a*v1+v2) = a*attn(q,k,v1)+attn(q,k,v2)
    v2 = torch.randn(2, 64 * 80, 128, generator=g).to(DEV)
    lhs = ctx.window_attn(q.to(DEV), k.to(DEV), 0.5 * v.to(DEV) + v2, 64, 80, 2, shift, impl=impl)
    rhs = 0.5 * out + ctx.window_attn(q.to(DEV), k.to(DEV), v2, 64, 80, 2, shift, impl=impl)
    assert rms(lhs, rhs) < (1e-5 if impl == 1 else 3e-3)


# ------------------------------------------------------------------------------------------- backbone instance norm
@pytest.mark.parametrize("shape", [(3, 64, 32, 40), (2, 96, 7, 9), (1, 5, 1, 3), (3, 128, 64, 80)])
def test_instance_norm_fused_modes(ctx, shape):
    """mnf_instance_norm_fwd vs the PyTorch ops of the reference backbone (models/gmflow/backbone.py:28-36): fp32
    round-off only; odd plane sizes take the scalar path."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g) * 3 + 0.5).to(DEV)
    res = torch.randn(*shape, generator=g).to(DEV)
    ref0 = F.instance_norm(x.double()).float()
    assert max_abs(ctx.instance_norm(x, 0), ref0) < 2e-5
    assert max_abs(ctx.instance_norm(x, 1), F.relu(ref0)) < 2e-5
    assert max_abs(ctx.instance_norm(x, 2, res), F.relu(res + F.relu(ref0))) < 2e-5
    # a constant plane: variance 0 -> output 0 (eps keeps it finite)
    c = torch.full(shape, 2.5, device=DEV)
    assert float(ctx.instance_norm(c, 0).abs().max()) < 1e-3


@pytest.mark.parametrize("h,w,splits,shift", [(100, 100, 2, True), (96, 128, 2, False), (96, 128, 4, True), (80, 120, 4, True),
                                              (80, 120, 2, False)])
def test_window_attn_blender_llff_shapes(ctx, h, w, splits, shift):
    """BASELINE configs[3] (Blender 800x800: 100x100 tokens, windows of 2500 keys = 19 full + 1 partial key tile), the LLFF /
    IBRNet size (768x1024: 96x128 tokens, attn_splits 2 and 4 -- configs/train_ibrnet.yaml:9) and the 640x960 video demo
    (80x120 tokens, attn_splits 4: windows of 600 keys -- configs/test_video_own.yaml:8): both kernels vs the CPU ORACLE
    (oracle/encoder_oracle.py, pinned against the reference) on batch item 0, and vs each other on the whole batch."""
    g = torch.Generator().manual_seed(h + w)
    qc, kc, vc = (torch.randn(2, h * w, 128, generator=g) for _ in range(3))
    q, k, v = qc.to(DEV), kc.to(DEV), vc.to(DEV)
    oracle = EO.window_attention(qc[:1], kc[:1], vc[:1], h, w, splits, shift)
    ref = ctx.window_attn(q, k, v, h, w, splits, shift, impl=1)
    got = ctx.window_attn(q, k, v, h, w, splits, shift, impl=2)
    nows = ctx.window_attn(q, k, v, h, w, splits, shift, impl=2, use_workspace=False)
    torch.cuda.synchronize()
    assert rms(ref[:1], oracle) < 2e-6 * max(1.0, float(oracle.std()))
    assert rms(got[:1], oracle) < 2e-3 * float(oracle.std()) and rms(nows[:1], oracle) < 2e-3 * float(oracle.std())
    assert max_abs(got[:1], oracle) < 2e-2
    assert rms(got, ref) < 2e-3 * float(ref.std()) and rms(nows, ref) < 2e-3 * float(ref.std())


@pytest.mark.parametrize("h,w,splits,shift,cross", [(64, 80, 2, False, False), (64, 80, 2, True, True), (100, 100, 2, True, True),
                                                    (8, 12, 2, True, True), (6, 10, 1, False, True), (80, 120, 4, True, False)])
def test_window_attn_fused_projection(ctx, h, w, splits, shift, cross):
    """mnf_window_attn_proj_fwd = attention(q_proj(source), k_proj(target), v_proj(target)) (models/gmflow/transformer.py:158-171)
    with the projections inside the operand-packing kernel: vs the CPU oracle (projections in fp32, oracle attention) on batch item
    0, and vs the unfused path of this library (fp32 projections by torch, then mnf_window_attn_fwd) on the whole batch.  Self
    (source is target) and cross attention; DTU, Blender (partial tiles), tiny, full-attention and 4 x 4 window shapes."""
    g = torch.Generator().manual_seed(h * w + splits)
    B = 2
    src = torch.randn(B, h * w, 128, generator=g)
    tgt = torch.randn(B, h * w, 128, generator=g) if cross else src
    wq, wk, wv = (torch.randn(128, 128, generator=g) / 128 ** 0.5 * gain for gain in (1.6, 1.6, 1.0))
    blob = ctx.window_attn_pack_proj(wq.to(DEV), wk.to(DEV), wv.to(DEV))
    got = ctx.window_attn_proj(src.to(DEV), tgt.to(DEV), blob, h, w, splits, shift)
    got2 = ctx.window_attn_proj(src.to(DEV), tgt.to(DEV), blob, h, w, splits, shift)
    torch.cuda.synchronize()
    assert torch.equal(got, got2)
    oracle = EO.window_attention(src[:1] @ wq.T, tgt[:1] @ wk.T, tgt[:1] @ wv.T, h, w, splits, shift)
    sd = float(oracle.std())
    assert rms(got[:1], oracle) < 3e-3 * sd and max_abs(got[:1], oracle) < 3e-2 * max(1.0, sd), (rms(got[:1], oracle), sd)
    unfused = ctx.window_attn(src.to(DEV) @ wq.to(DEV).T, tgt.to(DEV) @ wk.to(DEV).T, tgt.to(DEV) @ wv.to(DEV).T, h, w, splits, shift, impl=2)
    assert rms(got, unfused) < 3e-3 * float(unfused.std())
    # target_batch_roll: keys / values of item b from target[(b + 1) % B] == the same call on the rolled tensor, bit for bit
    rolled = ctx.window_attn_proj(src.to(DEV), tgt.to(DEV), blob, h, w, splits, shift, target_roll=1)
    explicit = ctx.window_attn_proj(src.to(DEV), torch.roll(tgt, -1, 0).to(DEV), blob, h, w, splits, shift)
    assert torch.equal(rolled, explicit) and (not cross or not torch.equal(rolled, got))


@pytest.mark.parametrize("rows", [1, 37, 5120 * 6])
def test_token_layernorm_fused_modes(ctx, rows):
    """mnf_token_layernorm_fwd vs nn.LayerNorm + the add / cat that follow it in TransformerLayer.forward
    (models/gmflow/transformer.py:173-185): fp32 arithmetic, tolerance 2e-5 (fp16 outputs: one fp16 rounding)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(rows)
    dev = ctx.device
    x = (torch.randn(rows, 128, generator=g) * 3 + 0.7).to(dev)
    src = torch.randn(rows, 128, generator=g).to(dev)
    w = (1 + 0.1 * torch.randn(128, generator=g)).to(dev)
    b = (0.1 * torch.randn(128, generator=g)).to(dev)
    ref = F.layer_norm(x.double(), (128,), w.double(), b.double(), 1e-5)
    assert max_abs(ctx.token_layernorm(x, w, b, 1e-5), ref.float()) < 2e-5
    assert max_abs(ctx.token_layernorm(x, w, b, 1e-5, residual=src), (src.double() + ref).float()) < 2e-5
    cat = ctx.token_layernorm(x, w, b, 1e-5, prefix=src)
    assert cat.dtype == torch.float16 and cat.shape == (rows, 256)
    assert torch.equal(cat[:, :128], src.half()) and max_abs(cat[:, 128:].float(), ref.float()) < 4e-3
    xh = x.half()                                        # fp16 input (the FFN output)
    refh = F.layer_norm(xh.double(), (128,), w.double(), b.double(), 1e-5)
    assert max_abs(ctx.token_layernorm(xh, w, b, 1e-5, residual=src), (src.double() + refh).float()) < 2e-5
    with pytest.raises(ValueError):
        ctx.token_layernorm(x, w, b, 1e-5, residual=src, prefix=src)


@pytest.mark.parametrize("ffn", [False, True])
@pytest.mark.parametrize("rows", [1, 37, 128, 300, 148 * 128 + 5, 5120 * 6])
def test_token_block(ctx, rows, ffn):
    """mnf_token_block_fwd (merge + LayerNorm [+ FFN + LayerNorm] + residual in one tcgen05 kernel) vs the oracle's restatement of
    models/gmflow/transformer.py:173-185 on the same inputs.  Two references: the oracle with every GEMM operand rounded to fp16
    where the kernel rounds it (sharp: 3e-4 RMS -- accumulation order, the GELU fit, rsqrt) and the plain fp32 oracle (the
    mixed-precision budget: 3e-3 RMS of O(1) outputs).  Ragged token counts, more tiles than SMs, DTU size (6 x 64 x 80 tokens)."""
    g = torch.Generator().manual_seed(rows + int(ffn))
    sd = {"merge.weight": torch.randn(128, 128, generator=g) / 128 ** 0.5,
          "norm1.weight": 1 + 0.2 * torch.randn(128, generator=g), "norm1.bias": 0.2 * torch.randn(128, generator=g),
          "mlp.0.weight": torch.randn(1024, 256, generator=g) / 256 ** 0.5 * 1.5,
          "mlp.2.weight": torch.randn(128, 1024, generator=g) / 1024 ** 0.5,
          "norm2.weight": 1 + 0.2 * torch.randn(128, generator=g), "norm2.bias": 0.2 * torch.randn(128, generator=g)}
    attn = torch.randn(rows, 128, generator=g) * 1.3
    src = torch.randn(rows, 128, generator=g)
    dev = ctx.device
    blob = ctx.token_block_pack(sd["merge.weight"].to(dev), sd["norm1.weight"].to(dev), sd["norm1.bias"].to(dev),
                                *([sd[k].to(dev) for k in ("mlp.0.weight", "mlp.2.weight", "norm2.weight", "norm2.bias")] if ffn else []))
    out = ctx.token_block(attn.to(dev), src.to(dev), blob, ffn)
    out2 = ctx.token_block(attn.to(dev), src.to(dev), blob, ffn)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)                                         # deterministic
    sub = slice(None) if rows <= 4096 else torch.randperm(rows, generator=g)[:4096]
    sd64 = {k: v.double() for k, v in sd.items()}
    ref_q = EO.post_attention(sd64, "", src[sub].double(), attn[sub].double(), ffn, quant=lambda t: t.half().double())
    ref = EO.post_attention(sd64, "", src[sub].double(), attn[sub].double(), ffn)
    got = out.cpu()[sub].double()
    assert bool(torch.isfinite(got).all())
    assert rms(got, ref_q) < 3e-4 and max_abs(got, ref_q) < 1e-2, (rms(got, ref_q), max_abs(got, ref_q))
    assert rms(got, ref) < 3e-3, rms(got, ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("shape", [(3, 64, 32, 40), (2, 96, 7, 9), (1, 8, 1, 3), (3, 128, 64, 80), (3, 64, 256, 320)])
def test_instance_norm_nhwc_modes(ctx, shape, dtype):
    """mnf_instance_norm_nhwc_fwd (channels-last activations of the backbone, fp32 or fp16) vs the PyTorch ops of the reference
    backbone (models/gmflow/backbone.py:28-36) in float64 on the same inputs: fp32 2e-5, fp16 one rounding of the output."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g) * 2.5 + 0.3).to(dtype).to(ctx.device).contiguous(memory_format=torch.channels_last)
    res = torch.randn(*shape, generator=g).to(dtype).to(ctx.device).contiguous(memory_format=torch.channels_last)
    ref0 = F.instance_norm(x.double())
    rel = 1.5e-3 if dtype == torch.float16 else 2e-5
    for mode, ref in ((0, ref0), (1, F.relu(ref0)), (2, F.relu(res.double() + F.relu(ref0)))):
        y = ctx.instance_norm_nhwc(x, mode, res if mode == 2 else None)
        assert y.dtype == dtype and y.shape == x.shape and y.is_contiguous(memory_format=torch.channels_last)
        assert max_abs(y.float(), ref.float()) < rel * max(1.0, float(ref.abs().max())), (mode, max_abs(y.float(), ref.float()))
    with pytest.raises(ValueError):
        ctx.instance_norm_nhwc(x.contiguous(), 0)           # not channels_last
