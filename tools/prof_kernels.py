#!/usr/bin/env python
"""Run the per-ray kernels in isolation on a synthetic DTU-sized scene (for ncu captures and quick timing).

    python tools/prof_kernels.py [--rays 81920] [--samples 64] [--reps 3] [--which gather,decoder,attn] [--impl 0]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from matchnerf_b200 import capi  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=81920)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--which", default="gather,decoder,attn")
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--random-feats", action="store_true", help="N(0,1) feature maps instead of real encoder output")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = capi.get_context(dev)
    H, W, S = 512, 640, args.samples
    feats, imgs, _ = synth.synthetic_scene(H, W, seed=1234)
    extr, intr, nf = synth.synthetic_cameras(H, W)
    ctx.load_decoder(synth.synthetic_decoder(0))
    if not args.random_feats:
        # feature maps produced by the encoder (synthetic weights) on random images: the decoder's data-dependent paths
        # (activation magnitudes, attention score ranges) then match what bench.py runs
        from matchnerf_b200.matchnerf import MatchNeRF
        from bench import make_opts
        m = MatchNeRF(make_opts(S, str(dev))).eval()
        m.feat_enc.load_state_dict(synth.synthetic_encoder(1))
        m.to(dev)
        with torch.no_grad():
            f = m.get_img_feat(imgs.to(dev))
        feats = [f[0].cpu(), f[1].cpu()]
        del m
    packed = ctx.pack_scene([feats[0][0].to(dev), feats[1][0].to(dev)], imgs[0].to(dev), extr[0, :3], intr[0, :3], nf[0, :3])
    sc = packed.c_scene(extr[0, 3, :3], intr[0, 3], nf[0, 3])
    cfg = capi.DecoderCfg()
    cfg.n_samples, cfg.raytrans_act, cfg.raytrans_posenc, cfg.density_maskfill = S, 0, 0, 0
    which = args.which.split(",")
    first = min(100 * W, H * W - args.rays)

    def timeit(name, fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        print(f"{name}: {ms:.3f} ms per launch ({args.rays} rays x {S} samples) -> {args.rays / ms * 1e3 / 1e6:.2f} M rays/s", flush=True)

    c32 = c16 = None
    if "gather" in which or "decoder" in which:
        timeit("gather_cossim", lambda: ctx.gather_cossim(sc, S, first_ray=first, n_rays=args.rays, want_f32=False, want_f16=True))
        c32, c16 = ctx.gather_cossim(sc, S, first_ray=first, n_rays=args.rays, want_f32=True, want_f16=True)
    if "decoder" in which:
        for impl in ([1, 2] if args.impl == 0 else [args.impl]):
            try:
                timeit(f"decoder_composite impl={impl}", lambda: ctx.decoder_composite(sc, cfg, cond_f32=c32, cond_f16=c16, first_ray=first,
                                                                                      n_rays=args.rays, impl=impl))
            except RuntimeError as e:
                print(f"decoder impl={impl}: {e}")
    if "encoder" in which:
        from matchnerf_b200.matchnerf import MatchNeRF
        sys.path.insert(0, ROOT)
        from bench import make_opts
        opt = make_opts(S, str(dev))
        model = MatchNeRF(opt).eval()
        model.feat_enc.load_state_dict(synth.synthetic_encoder(1))
        model.to(dev)
        enc = model.feat_enc
        images = imgs.to(dev)
        with torch.no_grad():
            timeit("encoder total (get_img_feat)", lambda: model.get_img_feat(images))
            x = enc.normalize_images(images).reshape(3, 3, H, W)
            timeit("  backbone (3 views)", lambda: enc.backbone(x))
            base = enc.backbone(x)
            f0 = torch.stack([base[0], base[0], base[1]])
            f1 = torch.stack([base[1], base[2], base[2]])
            timeit("  transformer (3 pairs, 6 blocks)", lambda: enc.transformer(f0, f1, 2))
            t0, t1 = enc.transformer(f0, f1, 2)
            timeit("  upsampler", lambda: enc.featup_net(torch.cat([t0, t1], 0)))
            layer = enc.transformer.layers[0].cross_attn_ffn
            xx = torch.cat([f0, f1], 0).flatten(2).transpose(1, 2).contiguous()
            timeit("    one cross_attn_ffn layer", lambda: layer(xx, xx, 64, 80, 2))
            timeit("    its q/k/v/merge projections + norms only", lambda: (layer.norm1(layer.merge(layer.q_proj(xx))), layer.k_proj(xx), layer.v_proj(xx)))
            timeit("    its FFN only", lambda: layer.norm2(layer.mlp(torch.cat([xx, xx], dim=-1))))
    if "attn" in which:
        g = torch.Generator().manual_seed(0)
        q, k, v = (torch.randn(6, 64 * 80, 128, generator=g).to(dev) for _ in range(3))
        for impl in ([1, 2] if args.impl == 0 else [args.impl]):
            for shift in (False, True):
                try:
                    timeit(f"window_attn impl={impl} shift={int(shift)}", lambda: ctx.window_attn(q, k, v, 64, 80, 2, shift, impl=impl))
                except RuntimeError as e:
                    print(f"attn impl={impl}: {e}")


if __name__ == "__main__":
    main()
