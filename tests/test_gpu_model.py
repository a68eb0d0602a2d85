"""GPU (B200): the reference-facing model interface end to end -- encoder (K-attn inside) + renderer -- against the
oracle and the reference goldens, plus full-image properties at BASELINE config-2 size."""
import os

import numpy as np
import pytest
import torch
import yaml

from matchnerf_b200.utils import AttrDict
from oracle import encoder_oracle as EO
from oracle import synth
from tests.helpers import load_npz, oracle_render, psnr, rms
from tests.test_host_cpu import make_opts

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build_model(S, **over):
    from matchnerf_b200.matchnerf import MatchNeRF
    opt = make_opts(**{"nerf.sample_intvs": S, **over})
    opt.device = DEV
    m = MatchNeRF(opt).eval()
    m.feat_enc.load_state_dict(synth.synthetic_encoder(1), strict=True)
    m.nerf_dec.load_state_dict(synth.synthetic_decoder(0), strict=True)
    return m.to(DEV), opt


def test_encoder_matches_reference_golden(golden_dir):
    z = load_npz(golden_dir, "encoder_64x96.npz")
    m, opt = build_model(16)
    with torch.no_grad():
        f8, f4 = m.get_img_feat(torch.from_numpy(z["images"]).to(DEV))
    scale = float(np.std(z["feat8"]))
    # TF32 cuDNN convolutions / cuBLAS GEMMs on the GPU vs the reference's fp32 CPU run: 10-bit-mantissa operands in ~40 chained
    # layers leave 1.9e-3 (NCHW algorithms) to 2.1e-3 (NHWC algorithms) relative RMS on this tiny 8 x 12 map (tools/r02_golden_err.py;
    # 1.7e-3 at DTU size) -- library arithmetic, not this repo's kernels, and the rendered image stays at 4.4e-4 rgb RMS from the
    # reference (bench.py parity; budget 2e-3).  The fp16-activation backbone (opt-in) measures 2.9e-3 here.
    assert rms(f8[0], z["feat8"]) < 2.5e-3 * scale, (rms(f8[0], z["feat8"]), scale)
    assert rms(f4[0][:, ::8], z["feat4_ch0mod8"]) < 2.5e-3 * scale


def test_encoder_token_path_equals_reference_shaped_path():
    """GMFlow._forward_tokens (activations kept in the token / channels-last layout, the other view addressed by a batch roll inside
    the attention kernel) vs the reference-shaped glue around the same kernels (NCHW stacks, cat + transpose, per-block cat of the
    swapped halves): identical bits, for all pairs and for a pair subset (the multi-GPU encoder split)."""
    from matchnerf_b200.gmflow import GMFlow
    m, opt = build_model(16)
    m.encoder_cuda_graph = False
    imgs = torch.rand(2, 3, 3, 64, 96, generator=torch.Generator().manual_seed(9)).to(DEV)
    enc = m.feat_enc
    with torch.no_grad():
        for pair_ids in (None, [2, 0]):
            assert enc._token_path_ok(imgs)
            a = enc(imgs=imgs, attn_splits_list=[2], keep_raw_feats=True, pair_ids=pair_ids)
            GMFlow.token_path = False
            try:
                b = enc(imgs=imgs, attn_splits_list=[2], keep_raw_feats=True, pair_ids=pair_ids)
            finally:
                GMFlow.token_path = True
            for k in ("aug_feat0s", "aug_feat1s"):
                assert len(a[k]) == len(b[k]) == 2
                for x, y in zip(a[k], b[k]):
                    assert x.shape == y.shape and torch.equal(x, y)


@pytest.mark.parametrize("local", [(0, 1), (1, 2)])
def test_forward_small_image_vs_oracle(local):
    """forward(mode='test') on a 64x96 triplet: encoder + full-image render, vs the CPU oracle chain; also with
    encoder.feature_sample_local_radius / _dilation = (1, 2) (models/gmflow/utils.py:136-162)."""
    H, W, S = 64, 96, 32
    m, opt = build_model(S, **{"encoder.feature_sample_local_radius": local[0], "encoder.feature_sample_local_dilation": local[1]})
    g = torch.Generator().manual_seed(4)
    images = torch.rand(1, 4, 3, H, W, generator=g)
    extr, intr, nf = synth.synthetic_cameras(H, W)
    batch = AttrDict(images=images.to(DEV), extrinsics=extr.to(DEV), intrinsics=intr.to(DEV), near_fars=nf.to(DEV))
    with torch.no_grad():
        out = m(batch, mode="test")
    assert out.rgb.shape == (1, H * W, 3) and out.depth.shape == (1, H * W, 1) and out.opacity.shape == (1, H * W, 1)
    feats = EO.encode_views(synth.synthetic_encoder(1), images[0, :3])
    idx = torch.arange(0, H * W, 7)
    o = oracle_render(synth.synthetic_decoder(0), [f[None] for f in feats], images[:, :3], extr, intr, nf, idx, S,
                      local_radius=local[0], local_dilation=local[1])
    e = (rms(out.rgb[0, idx], o[0]), rms(out.depth[0, idx], o[1]), rms(out.opacity[0, idx], o[2]))
    assert e[0] < 2e-3 and e[2] < 4e-3, e
    assert 0.02 < float(out.opacity.mean()) < 0.98


def test_train_mode_random_rays_and_slicing_invariance():
    H, W, S = 64, 96, 16
    m, opt = build_model(S, **{"nerf.rand_rays_train": 256, "nerf.rand_rays_test": 1000})
    g = torch.Generator().manual_seed(5)
    batch = AttrDict(images=torch.rand(1, 4, 3, H, W, generator=g).to(DEV))
    extr, intr, nf = synth.synthetic_cameras(H, W)
    batch.update(extrinsics=extr.to(DEV), intrinsics=intr.to(DEV), near_fars=nf.to(DEV))
    with torch.no_grad():
        out = m(AttrDict(batch), mode="train")
        assert out.ray_idx.shape == (256,) and out.rgb.shape == (1, 256, 3)
        full = m(AttrDict(batch), mode="test")
        m.render_chunk = 999                         # different slicing must not change any pixel
        full2 = m(AttrDict(batch), mode="test")
    if os.environ.get("MNF_GATHER_IMPL") == "4":     # tensor-core gather experiment: a ray's result depends on its 16-ray batch in the last bits
        assert rms(full.rgb, full2.rgb) < 1e-5 and rms(full.depth, full2.depth) < 1e-4
    else:
        assert torch.equal(full.rgb, full2.rgb) and torch.equal(full.depth, full2.depth)


def test_full_size_dtu_properties():
    """BASELINE config 2 size (512x640, S=64): properties that do not need the CPU oracle at full size --
    opacity in [0,1], rgb in [0,1], background compositing identity rgb_bg = rgb + (1 - opacity),
    and agreement of a strided subset with the oracle."""
    H, W, S = 512, 640, 64
    m, opt = build_model(S)
    feats, imgs, g = synth.synthetic_scene(H, W, seed=1234)
    extr, intr, nf = synth.synthetic_cameras(H, W)
    tgt = dict(extrinsics=extr[:, 3, :3].to(DEV), intrinsics=intr[:, 3].to(DEV), near_fars=nf[:, 3].to(DEV))
    ref = dict(extrinsics=extr[:, :3, :3].to(DEV), intrinsics=intr[:, :3].to(DEV), near_fars=nf[:, :3].to(DEV))
    fd = [f.to(DEV) for f in feats]
    with torch.no_grad():
        a = m.render_by_slices(opt, tgt, mode="test", ref_poses=ref, ref_images=imgs.to(DEV), ref_feats_list=fd)
        m.nerf_setbg_opaque = True
        b = m.render_by_slices(opt, tgt, mode="test", ref_poses=ref, ref_images=imgs.to(DEV), ref_feats_list=fd)
    assert a.rgb.shape == (1, H * W, 3)
    assert float(a.opacity.min()) >= 0 and float(a.opacity.max()) <= 1 + 1e-5
    assert float(a.rgb.min()) >= 0 and float(a.rgb.max()) <= 1 + 1e-5
    assert rms(b.rgb, a.rgb + (1 - a.opacity)) < 1e-6
    idx = torch.arange(1000, H * W, 4099)
    o = oracle_render(synth.synthetic_decoder(0), feats, imgs, extr, intr, nf, idx, S)
    assert rms(a.rgb[0, idx], o[0]) < 2e-3


def test_video_mode_renders_every_frame_of_the_path():
    """forward(render_video=True) (models/matchnerf.py:42-71): one encoder pass, one full render per path pose, frames
    moved to the host and concatenated on dim 0; each frame equals a stand-alone render at that pose."""
    H, W, S = 64, 96, 16
    m, opt = build_model(S, **{"nerf.rand_rays_test": 2000})
    opt.nerf["video_n_frames"] = 6
    g = torch.Generator().manual_seed(8)
    images = torch.rand(1, 4, 3, H, W, generator=g)
    extr, intr, nf = synth.synthetic_cameras(H, W)
    batch = AttrDict(images=images.to(DEV), extrinsics=extr.to(DEV), intrinsics=intr.to(DEV), near_fars=nf.to(DEV))
    with torch.no_grad():
        out = m(AttrDict(batch), mode="test", render_video=True, render_path_mode="interpolate")
        assert out.rgb.shape == (6, H * W, 3) and out.depth.shape == (6, H * W, 1) and not out.rgb.is_cuda
        tgt, ref = m.extract_poses(batch)
        frames = m.get_video_rendering_path(tgt, ref, "interpolate", 6, batch)
        feats = m.get_img_feat(batch["images"][:, :3])
        for f in (0, 3, 5):
            one = m.render_by_slices(opt, frames[f], mode="test", ref_poses=ref, ref_images=batch["images"][:, :3], ref_feats_list=feats)
            assert rms(one.rgb[0], out.rgb[f]) < 1e-5 and rms(one.opacity[0], out.opacity[f]) < 1e-5
    assert 0.02 < float(out.opacity.mean()) < 0.98
    with pytest.raises(AssertionError):
        m(AttrDict(batch), mode="train", render_video=True)


def test_blender_size_render_properties():
    """BASELINE configs[3] size (Blender 800x800, wide baseline, white background): full-image render from given feature
    maps (100x100 and 200x200, i.e. not L2-resident in fp32), background identity, and a strided subset vs the oracle."""
    H, W, S = 800, 800, 32
    m, opt = build_model(S)
    g = torch.Generator().manual_seed(21)
    feats = [torch.randn(1, 3, 256, H // 8, W // 8, generator=g), torch.randn(1, 3, 256, H // 4, W // 4, generator=g)]
    imgs = torch.rand(1, 3, 3, H, W, generator=g)
    extr, intr, nf = synth.synthetic_cameras(H, W, baseline_deg=30.0, near=2.0, far=6.0)
    tgt = dict(extrinsics=extr[:, 3, :3].to(DEV), intrinsics=intr[:, 3].to(DEV), near_fars=nf[:, 3].to(DEV))
    ref = dict(extrinsics=extr[:, :3, :3].to(DEV), intrinsics=intr[:, :3].to(DEV), near_fars=nf[:, :3].to(DEV))
    fd = [f.to(DEV) for f in feats]
    with torch.no_grad():
        a = m.render_by_slices(opt, tgt, mode="test", ref_poses=ref, ref_images=imgs.to(DEV), ref_feats_list=fd)
        m.nerf_setbg_opaque = True
        b = m.render_by_slices(opt, tgt, mode="test", ref_poses=ref, ref_images=imgs.to(DEV), ref_feats_list=fd)
    assert a.rgb.shape == (1, H * W, 3)
    assert float(a.opacity.min()) >= 0 and float(a.opacity.max()) <= 1 + 1e-5
    assert rms(b.rgb, a.rgb + (1 - a.opacity)) < 1e-6
    idx = torch.arange(777, H * W, 9973)
    o = oracle_render(synth.synthetic_decoder(0), feats, imgs, extr, intr, nf, idx, S)
    assert rms(a.rgb[0, idx], o[0]) < 2e-3 and rms(a.opacity[0, idx], o[2]) < 4e-3
    assert 0.02 < float(a.opacity.mean()) < 0.98


def test_batch_of_two_scenes_equals_two_single_forwards():
    """Batch schema with B = 2 (models/matchnerf.py:32-73 loops nothing: every tensor carries the batch dim): two different
    scenes and camera sets in one forward equal the two single-scene forwards; train mode shares one ray_idx across the batch."""
    H, W, S = 64, 96, 16
    m, opt = build_model(S, **{"nerf.rand_rays_train": 512})
    g = torch.Generator().manual_seed(12)
    images = torch.rand(2, 4, 3, H, W, generator=g)
    e0, i0, n0 = synth.synthetic_cameras(H, W)
    e1, i1, n1 = synth.synthetic_cameras(H, W, baseline_deg=18.0, near=2.5, far=5.0)
    extr, intr, nf = torch.cat([e0, e1]), torch.cat([i0, i1]), torch.cat([n0, n1])
    both = AttrDict(images=images.to(DEV), extrinsics=extr.to(DEV), intrinsics=intr.to(DEV), near_fars=nf.to(DEV))
    with torch.no_grad():
        out = m(AttrDict(both), mode="test")
        assert out.rgb.shape == (2, H * W, 3) and out.depth.shape == (2, H * W, 1) and out.opacity.shape == (2, H * W, 1)
        for b in range(2):
            one = m(AttrDict(images=images[b:b + 1].to(DEV), extrinsics=extr[b:b + 1].to(DEV), intrinsics=intr[b:b + 1].to(DEV),
                             near_fars=nf[b:b + 1].to(DEV)), mode="test")
            # not bit-equal: the encoder's TF32 GEMMs pick other tilings at twice the batch (feature maps move ~1e-3 relative)
            assert rms(one.rgb[0], out.rgb[b]) < 1e-3 and rms(one.depth[0], out.depth[b]) < 1e-2 and rms(one.opacity[0], out.opacity[b]) < 2e-3
        tr = m(AttrDict(both), mode="train")
        assert tr.ray_idx.shape == (256,) and tr.rgb.shape == (2, 256, 3)
    assert rms(out.rgb[0], out.rgb[1]) > 2e-2                             # the two scenes really differ (a mix-up would show)
