#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Dev-container only (the reference does not travel to the GPU box).  For every case the
script (1) runs the reference, (2) runs this repo's oracle restatement on the same inputs
and asserts agreement, (3) stores inputs (or the recipe to regenerate them) and the
reference outputs.  It also re-derives the SURVEY.md 8(c) known-answer numbers (reference
default init, seed 0) to pin the harness itself.

    python oracle/make_golden.py            # writes tests/golden/
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("MATCHNERF_REFERENCE", "/root/reference")


# ---- import shim: stubs for the absent third-party modules on the reference's import chain
from oracle.reference_shim import EasyDict, install_shim as _install_shim, reference_options  # noqa: E402


def install_shim():
    _install_shim(REF)


def ref_options(S: int, **over):
    return reference_options(S, "cpu", REF, **over)


def main():
    install_shim()
    from models.matchnerf import MatchNeRF                      # the reference
    from oracle import encoder_oracle as EO
    from oracle import render_oracle as RO
    from oracle import synth

    out_dir = os.environ.get("MATCHNERF_GOLDEN_OUT") or os.path.join(ROOT, "tests", "golden")   # override: reproducibility check
    os.makedirs(out_dir, exist_ok=True)
    torch.set_grad_enabled(False)
    report = []

    def rms(a, b):
        return float(((a - b) ** 2).mean().sqrt())

    # ------------------------------------------------------------------ known answers of SURVEY 8(c)
    for S, exp in ((64, dict(rgb=0.1042879, depth=0.6255180, op=0.1885468)),
                   (128, dict(rgb=0.1714871, depth=1.0115364, op=0.3102203))):
        torch.manual_seed(0)
        opt = ref_options(S)
        model = MatchNeRF(opt).eval()
        feats, imgs, g = synth.synthetic_scene(512, 640, seed=1234)
        extr, intr, nf = synth.synthetic_cameras(512, 640)
        batch = EasyDict(images=None, extrinsics=extr, intrinsics=intr, near_fars=nf)
        ray_idx = torch.randperm(512 * 640, generator=g)[:1024]
        assert ray_idx[:4].tolist() == [253665, 137494, 291312, 108696]
        tgt, ref = model.extract_poses(batch)
        out = model.render(opt, tgt, ray_idx=ray_idx, mode="test", ref_poses=ref, ref_images=imgs, ref_feats_list=feats)
        got = dict(rgb=float(out.rgb.mean()), depth=float(out.depth.mean()), op=float(out.opacity.mean()))
        for k in exp:
            assert abs(got[k] - exp[k]) < 2e-6, (S, k, got[k], exp[k])
        # oracle vs reference on the same (reference-initialised) weights
        dec = {k: v.clone() for k, v in model.nerf_dec.state_dict().items()}
        o_rgb, o_depth, o_op = RO.render_rays(dec, RO.to_channels_last(feats), imgs[0].permute(0, 2, 3, 1).contiguous(),
                                              extr[0, :3, :3], intr[0, :3], nf[0, :3], extr[0, 3, :3], intr[0, 3], nf[0, 3],
                                              ray_idx, S)
        e = (rms(o_rgb, out.rgb[0]), rms(o_depth, out.depth[0]), rms(o_op, out.opacity[0]))
        assert max(e) < 2e-5, e
        report.append(f"known-answer S={S}: reference means {got}  oracle-vs-ref RMS {e}")
        np.savez_compressed(os.path.join(out_dir, f"config1_refinit_S{S}.npz"),
                            note="SURVEY 8c recipe; reference default init seed 0; inputs regenerate from oracle.synth",
                            ray_idx=ray_idx.numpy(), rgb=out.rgb[0].numpy(), depth=out.depth[0].numpy(),
                            opacity=out.opacity[0].numpy(),
                            **{"dec." + k: v.numpy() for k, v in dec.items()})

    # ------------------------------------------------------------------ config 1 with synthetic (non-zero-bias) weights
    for S in (64, 128):
        opt = ref_options(S)
        model = MatchNeRF(opt).eval()
        dec = synth.synthetic_decoder(seed=0)
        model.nerf_dec.load_state_dict(dec, strict=True)
        feats, imgs, g = synth.synthetic_scene(512, 640, seed=1234)
        extr, intr, nf = synth.synthetic_cameras(512, 640)
        batch = EasyDict(images=None, extrinsics=extr, intrinsics=intr, near_fars=nf)
        ray_idx = torch.randperm(512 * 640, generator=g)[:1024]
        tgt, ref = model.extract_poses(batch)
        out = model.render(opt, tgt, ray_idx=ray_idx, mode="test", ref_poses=ref, ref_images=imgs, ref_feats_list=feats)
        assert 0.05 < float(out.opacity.mean()) < 0.95, float(out.opacity.mean())
        o_rgb, o_depth, o_op = RO.render_rays(dec, RO.to_channels_last(feats), imgs[0].permute(0, 2, 3, 1).contiguous(),
                                              extr[0, :3, :3], intr[0, :3], nf[0, :3], extr[0, 3, :3], intr[0, 3], nf[0, 3],
                                              ray_idx, S)
        e = (rms(o_rgb, out.rgb[0]), rms(o_depth, out.depth[0]), rms(o_op, out.opacity[0]))
        assert max(e) < 2e-5, e
        report.append(f"config1 synth S={S}: opacity mean {float(out.opacity.mean()):.4f} rgb mean {float(out.rgb.mean()):.4f}; oracle-vs-ref RMS {e}")
        np.savez_compressed(os.path.join(out_dir, f"config1_synth_S{S}.npz"),
                            note="512x640, synthetic_scene(seed 1234), synthetic_cameras, synthetic_decoder(seed 0)",
                            ray_idx=ray_idx.numpy(), rgb=out.rgb[0].numpy(), depth=out.depth[0].numpy(),
                            opacity=out.opacity[0].numpy())

    # ------------------------------------------------------------------ small fully-stored render cases (option coverage)
    IBR = {"decoder.raytrans_act": "ELU", "decoder.density_maskfill": True, "decoder.raytrans_posenc": True}
    small_cases = [
        dict(name="small_base", S=16, over={}, bg=False, stratified=False),
        # ELU + raytrans_posenc + density_maskfill + white background: S = 32 runs on the tcgen05 kernel, S = 24 only on the fp32 one
        dict(name="small_elu_maskfill_posenc_bg", S=32, over=dict(IBR), bg=True, stratified=False),
        dict(name="small_s24_elu_maskfill_posenc_bg", S=24, over=dict(IBR), bg=True, stratified=False),
        dict(name="small_wide_baseline", S=32, over={}, bg=False, stratified=False, baseline=35.0),
        # configs/demo_own.yaml:10-15 (ELU + posenc + maskfill, S = 128) and configs/test_video_own.yaml:10-15 (same, S = 256);
        # wide baseline so that density_maskfill has samples no view sees
        dict(name="demo_own_S128", S=128, over=dict(IBR), bg=False, stratified=False, baseline=30.0, n_rays=96, gain=0.5),
        dict(name="video_own_S256", S=256, over=dict(IBR), bg=False, stratified=False, baseline=30.0, n_rays=48, gain=0.25),
        # encoder.feature_sample_local_radius > 0 (models/gmflow/utils.py:136-162; no shipped config): mean over 3 x 3 dilated taps
        dict(name="small_local_radius1", S=16, over={"encoder.feature_sample_local_radius": 1, "encoder.feature_sample_local_dilation": 2},
             bg=False, stratified=False),
    ]
    for case in small_cases:
        H, W, S = 32, 40, case["S"]
        opt = ref_options(S, **case["over"])
        model = MatchNeRF(opt).eval()
        dec = synth.synthetic_decoder(seed=0, density_gain=case.get('gain', 4.0))
        model.nerf_dec.load_state_dict(dec, strict=True)
        model.nerf_setbg_opaque = case["bg"]
        feats, imgs, g = synth.synthetic_scene(H, W, seed=77)
        extr, intr, nf = synth.synthetic_cameras(H, W, baseline_deg=case.get("baseline", 10.0))
        batch = EasyDict(images=None, extrinsics=extr, intrinsics=intr, near_fars=nf)
        ray_idx = torch.randperm(H * W, generator=g)[:case.get("n_rays", 200)]
        tgt, ref = model.extract_poses(batch)
        out = model.render(opt, tgt, ray_idx=ray_idx, mode="test", ref_poses=ref, ref_images=imgs, ref_feats_list=feats)
        # reference intermediate: conditioning vector of the same points
        centre, ray = RO.cast_rays(H, W, extr[0, 3, :3], intr[0, 3], ray_idx)
        o = RO.render_rays(dec, RO.to_channels_last(feats), imgs[0].permute(0, 2, 3, 1).contiguous(),
                           extr[0, :3, :3], intr[0, :3], nf[0, :3], extr[0, 3, :3], intr[0, 3], nf[0, 3], ray_idx, S,
                           setbg_opaque=case["bg"], raytrans_act=opt.decoder.raytrans_act,
                           raytrans_posenc=opt.decoder.raytrans_posenc, density_maskfill=opt.decoder.density_maskfill,
                           return_aux=True, local_radius=int(opt.encoder.feature_sample_local_radius),
                           local_dilation=int(opt.encoder.feature_sample_local_dilation))
        cond_ref = model.query_cond_info(o[3]["pts"].reshape(1, -1, S, 3), ref, imgs, feats)
        cond_ref = torch.cat([cond_ref["feat_info"], cond_ref["color_info"], cond_ref["mask_info"]], -1)[0].reshape(-1, 22)
        e = (rms(o[0], out.rgb[0]), rms(o[1], out.depth[0]), rms(o[2], out.opacity[0]), rms(o[3]["cond"], cond_ref))
        assert max(e) < 2e-5, (case["name"], e)
        assert 0.05 < float(out.opacity.mean()) < 0.95, (case["name"], float(out.opacity.mean()))
        report.append(f"{case['name']}: opacity mean {float(out.opacity.mean()):.4f}; oracle-vs-ref RMS {e}")
        np.savez_compressed(os.path.join(out_dir, case["name"] + ".npz"),
                            H=H, W=W, S=S, setbg_opaque=case["bg"],
                            raytrans_act=str(opt.decoder.raytrans_act), raytrans_posenc=bool(opt.decoder.raytrans_posenc),
                            density_maskfill=bool(opt.decoder.density_maskfill),
                            local_radius=int(opt.encoder.feature_sample_local_radius), local_dilation=int(opt.encoder.feature_sample_local_dilation),
                            feat8=feats[0].numpy(), feat4=feats[1].numpy(), images=imgs.numpy(),
                            extrinsics=extr.numpy(), intrinsics=intr.numpy(), near_fars=nf.numpy(),
                            ray_idx=ray_idx.numpy(), cond=cond_ref.numpy(),
                            rgb=out.rgb[0].numpy(), depth=out.depth[0].numpy(), opacity=out.opacity[0].numpy(),
                            **{"dec." + k: v.numpy() for k, v in dec.items()})

    # ------------------------------------------------------------------ window attention + encoder
    from models.gmflow.transformer import (generate_shift_window_attn_mask, single_head_full_attention,
                                           single_head_split_window_attention)
    g = torch.Generator().manual_seed(5)
    for (B, h, w, splits, shift) in ((2, 8, 12, 2, False), (2, 8, 12, 2, True), (1, 12, 16, 4, True), (2, 6, 10, 1, False)):
        q, k, v = (torch.randn(B, h * w, 128, generator=g) for _ in range(3))
        if splits == 1:
            ref_o = single_head_full_attention(q, k, v)
        else:
            mask = generate_shift_window_attn_mask((h, w), h // splits, w // splits, h // splits // 2, w // splits // 2,
                                                   device=torch.device("cpu")) if shift else None
            ref_o = single_head_split_window_attention(q, k, v, num_splits=splits, with_shift=shift, h=h, w=w, attn_mask=mask)
        my_o = EO.window_attention(q, k, v, h, w, splits, shift)
        e = rms(my_o, ref_o)
        assert e < 2e-6, (h, w, splits, shift, e)
        report.append(f"window_attn B{B} {h}x{w} splits{splits} shift{int(shift)}: oracle-vs-ref RMS {e:.2e}")
        np.savez_compressed(os.path.join(out_dir, f"window_attn_{h}x{w}_k{splits}_s{int(shift)}.npz"),
                            q=q.numpy(), k=k.numpy(), v=v.numpy(), h=h, w=w, num_splits=splits, with_shift=shift,
                            out=ref_o.numpy())

    H, W = 64, 96
    opt = ref_options(16)
    model = MatchNeRF(opt).eval()
    enc_sd = synth.synthetic_encoder(seed=1)
    model.feat_enc.load_state_dict(enc_sd, strict=True)
    n_par = sum(p.numel() for p in model.feat_enc.parameters())
    assert n_par == 4642208, n_par
    g = torch.Generator().manual_seed(9)
    imgs = torch.rand(1, 3, 3, H, W, generator=g)
    ref_feats = model.get_img_feat(imgs)
    my_feats = EO.encode_views(enc_sd, imgs[0])
    e = (rms(my_feats[0], ref_feats[0][0]), rms(my_feats[1], ref_feats[1][0]))
    scale = float(ref_feats[0].std())
    assert max(e) < 2e-5 * max(1.0, scale), (e, scale)
    report.append(f"encoder 64x96: feature std {scale:.3f}; oracle-vs-ref RMS {e}")
    np.savez_compressed(os.path.join(out_dir, "encoder_64x96.npz"),
                        note="synthetic_encoder(seed 1); images = rand(1,3,3,64,96, generator seed 9)",
                        images=imgs.numpy(), feat8=ref_feats[0][0].numpy(),
                        feat4_ch0mod8=ref_feats[1][0][:, ::8].contiguous().numpy())

    report.append(video_path_goldens(out_dir))

    with open(os.path.join(out_dir, "REPORT.txt"), "w") as f:
        f.write("generated by oracle/make_golden.py against the reference at %s (torch %s)\n" % (REF, torch.__version__))
        f.write("\n".join(report) + "\n")
    print("\n".join(report))


def video_path_goldens(out_dir: str) -> str:
    """Camera paths of the video mode (misc/camera.py:382-468) from the unmodified reference, for
    tests/test_host_cpu.py::test_video_paths_match_reference (host-side numpy, no GPU involved)."""
    from misc import camera as refcam                           # the reference
    from scipy.spatial.transform import Rotation
    from oracle import synth
    extr, _, _ = synth.synthetic_cameras(64, 96)
    w2c = torch.eye(4)[None].repeat(3, 1, 1)
    w2c[:, :3] = extr[0, :3, :3]
    c2w = torch.linalg.inv(w2c.double()).float().numpy()
    g = {"interp_in": c2w}
    for n in (6, 30):
        g[f"interp_{n}"] = refcam.get_interpolate_render_path(c2w[:, :3], n)
    rng = np.random.default_rng(0)
    eul = np.array([[170., 5., -175.], [-175., -3., 178.], [160., 10., 170.]])      # Euler angles straddling +-180 degrees
    c2w2 = np.tile(np.eye(4), (3, 1, 1))
    c2w2[:, :3, :3] = Rotation.from_euler("xyz", eul, degrees=True).as_matrix()
    c2w2[:, :3, 3] = rng.normal(size=(3, 3))
    g["interp_wrap_in"], g["interp_wrap_9"] = c2w2, refcam.get_interpolate_render_path(c2w2[:, :3], 9)
    M = 12
    c2wa = np.tile(np.eye(4), (M, 1, 1))
    c2wa[:, :3, :3] = Rotation.from_euler("xyz", rng.normal(scale=8, size=(M, 3)), degrees=True).as_matrix()
    c2wa[:, :3, 3] = rng.normal(scale=0.5, size=(M, 3))
    g["spiral_in"], g["spiral_10"] = c2wa, refcam.get_spiral_render_path(c2wa[:, :3], [2.0, 6.0], rads_scale=0.1, N_views=10)
    np.savez_compressed(os.path.join(out_dir, "video_paths.npz"), **g)
    return "video paths: interpolate (6, 30, wrap-around 9 frames) and spiral (10 frames) from misc/camera.py"


if __name__ == "__main__":
    main()
