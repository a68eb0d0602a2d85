"""CPU: host-side mirror of the reference interface -- module tree / state_dict compatibility, option plumbing,
feature regrouping, and the guarantee that nothing silently runs on the CPU."""
import os

import numpy as np

import pytest
import torch
import yaml

from matchnerf_b200.utils import AttrDict, get_opt
from oracle import synth

BASE_OPTS = """
n_src_views: 3
encoder: {attn_splits_list: [2], cos_n_group: [2, 8], num_transformer_layers: 6, feature_upsampler: network,
          upsample_factor: 2, wo_self_attn: false, feature_sample_local_radius: 0, feature_sample_local_dilation: 1}
decoder: {net_width: 128, net_depth: 6, skip: [4], posenc: {L_3D: 10, L_view: 0}, raytrans_posenc: false,
          density_maskfill: false, raytrans_act: ReLU}
nerf: {legacy_coord: true, wo_render_interval: true, view_dep: true, depth: {param: metric}, sample_intvs: 128,
       sample_stratified: true, rand_rays_test: 20480, rand_rays_train: 1024}
"""


def make_opts(**over):
    opt = AttrDict(yaml.safe_load(BASE_OPTS))
    opt.device = "cpu"
    for k, v in over.items():
        node = opt
        ks = k.split(".")
        for kk in ks[:-1]:
            node = node[kk]
        node[ks[-1]] = v
    return opt


def test_attrdict():
    d = AttrDict(a=1, b=dict(c=[dict(x=2)]))
    assert d.a == 1 and d.b.c[0].x == 2 and d["b"]["c"][0]["x"] == 2
    d.update(dict(e=dict(f=3)))
    assert d.e.f == 3 and get_opt(d, "e.f") == 3 and get_opt(d, "e.g.h", 7) == 7


def test_module_tree_matches_reference_state_dicts():
    from matchnerf_b200.matchnerf import MatchNeRF, models_dict
    assert models_dict["matchnerf"] is MatchNeRF
    m = MatchNeRF(make_opts())
    enc_shapes, dec_shapes = synth.encoder_param_shapes(), synth.decoder_param_shapes()
    enc_sd, dec_sd = m.feat_enc.state_dict(), m.nerf_dec.state_dict()
    assert list(dec_sd.keys()) == list(dec_shapes.keys())
    assert {k: tuple(v.shape) for k, v in dec_sd.items()} == dict(dec_shapes)
    assert set(enc_sd.keys()) == set(enc_shapes.keys())
    assert {k: tuple(v.shape) for k, v in enc_sd.items()} == dict(enc_shapes)
    # strict loading of reference-shaped checkpoints, per top-level child as misc/utils.py:183-205 does
    m.feat_enc.load_state_dict(synth.synthetic_encoder(1), strict=True)
    m.nerf_dec.load_state_dict(synth.synthetic_decoder(0), strict=True)
    assert {n for n, _ in m.named_children()} == {"feat_enc", "nerf_dec"}
    assert {n for n, _ in m.feat_enc.named_children()} == {"backbone", "transformer", "featup_net"}
    assert sum(p.numel() for p in m.feat_enc.parameters()) == 4642208
    assert sum(p.numel() for p in m.nerf_dec.parameters()) == 130324
    assert m.nerf_setbg_opaque is False and m.n_src_views == 3
    assert MatchNeRF.render_rays is MatchNeRF.render


def test_unsupported_architectures_are_rejected():
    from matchnerf_b200.matchnerf import MatchNeRF
    with pytest.raises(NotImplementedError):
        MatchNeRF(make_opts(**{"decoder.net_width": 256}))
    with pytest.raises(NotImplementedError):
        MatchNeRF(make_opts(**{"encoder.feature_sample_local_radius": 9}))
    with pytest.raises(NotImplementedError):
        MatchNeRF(make_opts(**{"nerf.legacy_coord": False}))


def test_decoder_cfg_from_options():
    from matchnerf_b200.matchnerf import MatchNeRF
    m = MatchNeRF(make_opts(**{"decoder.raytrans_act": "ELU", "decoder.density_maskfill": True, "nerf.sample_intvs": 64}))
    cfg = m.nerf_dec.decoder_cfg(m.opts)
    assert (cfg.n_samples, cfg.raytrans_act, cfg.raytrans_posenc, cfg.density_maskfill) == (64, 1, 0, 1)


def test_no_cpu_execution_path():
    """The hot path must fail loudly off-GPU instead of falling back to PyTorch/CPU math."""
    from matchnerf_b200.gmflow import window_attention
    from matchnerf_b200.matchnerf import MatchNeRF
    q = torch.randn(1, 16, 128)
    with pytest.raises(RuntimeError):
        window_attention(q, q, q, 4, 4, 2, False)
    m = MatchNeRF(make_opts()).eval()
    extr, intr, nf = synth.synthetic_cameras(32, 48)
    batch = AttrDict(images=torch.rand(1, 4, 3, 32, 48), extrinsics=extr, intrinsics=intr, near_fars=nf)
    with torch.no_grad(), pytest.raises(RuntimeError):
        m(batch, mode="test")


def test_feature_regrouping_matches_reference_convention():
    """get_img_feat: view i = concat of its features from each pair it is in (models/matchnerf.py:192-205)."""
    from matchnerf_b200.matchnerf import MatchNeRF
    m = MatchNeRF(make_opts())
    P, h, w = 3, 2, 3
    f0 = torch.arange(P, dtype=torch.float32).view(1, P, 1, 1, 1).expand(1, P, 128, h, w) + 10     # pair p, first member
    f1 = torch.arange(P, dtype=torch.float32).view(1, P, 1, 1, 1).expand(1, P, 128, h, w) + 20     # pair p, second member

    class FakeEnc(torch.nn.Module):
        def forward(self, imgs, **kw):
            return {"aug_feat0s": [f0], "aug_feat1s": [f1]}

    m.feat_enc = FakeEnc()
    out = m.get_img_feat(torch.zeros(1, 3, 3, 16, 24))[0]
    # pairs (0,1) (0,2) (1,2): view0 = [f0[0], f0[1]], view1 = [f1[0], f0[2]], view2 = [f1[1], f1[2]]
    halves = out[0, :, ::128, 0, 0]
    assert halves.tolist() == [[10.0, 11.0], [20.0, 12.0], [21.0, 22.0]]


def test_sine_position_matches_oracle():
    from matchnerf_b200.gmflow import sine_position
    from oracle.encoder_oracle import sine_position as ref
    assert torch.allclose(sine_position(4, 6, 128, "cpu"), ref(4, 6, 128), atol=1e-6)


def test_video_paths_match_reference(golden_dir):
    """camera_paths restates misc/camera.py:382-468; goldens come from the unmodified reference (oracle/make_golden.py)."""
    from matchnerf_b200 import camera_paths as cp
    from tests.helpers import load_npz
    z = load_npz(golden_dir, "video_paths.npz")
    for n in (6, 30):
        got = cp.interpolate_path(z["interp_in"][:, :3], n)
        assert got.shape == (3 * (n // 3), 4, 4) and np.abs(got - z[f"interp_{n}"]).max() < 1e-9
    assert np.abs(cp.interpolate_path(z["interp_wrap_in"][:, :3], 9) - z["interp_wrap_9"]).max() < 1e-9   # +-180 degree unwrap
    assert np.abs(cp.spiral_path(z["spiral_in"][:, :3], [2.0, 6.0], 0.1, 10) - z["spiral_10"]).max() < 1e-9


def test_video_rendering_path_poses():
    """get_video_rendering_path (models/matchnerf.py:295-325): n_frames pose dicts; the first interpolated frame is
    source view 0, frame n/3 is source view 1; spiral mode needs batch['c2ws_all']."""
    from matchnerf_b200.matchnerf import MatchNeRF
    m = MatchNeRF(make_opts()).eval()
    extr, intr, nf = synth.synthetic_cameras(32, 48)
    batch = AttrDict(images=torch.rand(1, 4, 3, 32, 48), extrinsics=extr, intrinsics=intr, near_fars=nf)
    tgt, ref = m.extract_poses(batch)
    frames = m.get_video_rendering_path(tgt, ref, "interpolate", n_frames=6, batch=batch)
    assert len(frames) == 6 and frames[0]["extrinsics"].shape == (1, 3, 4) and frames[0]["intrinsics"].shape == (1, 3, 3)
    assert torch.allclose(frames[0]["extrinsics"][0], extr[0, 0, :3], atol=1e-5)
    assert torch.allclose(frames[2]["extrinsics"][0], extr[0, 1, :3], atol=1e-5)
    assert torch.equal(frames[3]["near_fars"], tgt["near_fars"])
    with pytest.raises(Exception):
        m.get_video_rendering_path(tgt, ref, "zigzag", n_frames=6, batch=batch)
    sq = torch.eye(4).repeat(4, 1, 1)
    sq[:, :3] = extr[0, :, :3]
    batch["c2ws_all"] = torch.linalg.inv(sq)[None]
    sp = m.get_video_rendering_path(tgt, ref, "spiral", n_frames=5, batch=batch)
    assert len(sp) == 5 and torch.isfinite(sp[4]["extrinsics"]).all()
