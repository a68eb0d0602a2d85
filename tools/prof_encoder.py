#!/usr/bin/env python
"""Time variants of the encoder's library-call part (everything around K-attn) on a DTU-sized triplet, and report how
far each variant's feature maps move from the default path.

    python tools/prof_encoder.py [--reps 10]

Variants: default (NCHW, eager) | channels_last convolutions | CUDA-graph replay of the default | both.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import synth  # noqa: E402


def timeit(name, fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms:.3f} ms", flush=True)
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    from bench import make_opts
    from matchnerf_b200.matchnerf import MatchNeRF
    dev = torch.device("cuda", 0)
    H, W = 512, 640
    m = MatchNeRF(make_opts(64, str(dev))).eval()
    m.feat_enc.load_state_dict(synth.synthetic_encoder(1))
    m.to(dev)
    g = torch.Generator().manual_seed(7)
    images = torch.rand(1, 3, 3, H, W, generator=g).to(dev)
    enc = m.feat_enc

    def rel(a, b):
        return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())

    with torch.no_grad():
        ref = [f.clone() for f in m.get_img_feat(images)]
        timeit("encoder default (CUDA graph replay inside get_img_feat)", lambda: m.get_img_feat(images), args.reps)
        m.encoder_cuda_graph = False
        timeit("encoder eager", lambda: m.get_img_feat(images), args.reps)
        for mode in ("fp32",):
            enc.matmul_precision = mode
            f = m.get_img_feat(images)
            print(f"  precision {mode}: rel diff vs tf32 default {rel(f[0], ref[0]):.2e} / {rel(f[1], ref[1]):.2e}")
            timeit(f"encoder {mode}", lambda: m.get_img_feat(images), args.reps)
        enc.matmul_precision = "tf32"
        from matchnerf_b200.gmflow import TransformerLayer
        for dt in (None, torch.float16):
            TransformerLayer.ffn_dtype = dt
            f = m.get_img_feat(images)
            print(f"  ffn_dtype {dt}: rel diff vs default {rel(f[0], ref[0]):.2e} / {rel(f[1], ref[1]):.2e}")
            timeit(f"encoder ffn_dtype={dt}", lambda: m.get_img_feat(images), args.reps)
        # fp32 reference for the error budget of both variants
        enc.matmul_precision = "fp32"
        TransformerLayer.ffn_dtype = None
        f32 = [x.clone() for x in m.get_img_feat(images)]
        enc.matmul_precision = "tf32"
        for dt in (None, torch.float16):
            TransformerLayer.ffn_dtype = dt
            f = m.get_img_feat(images)
            print(f"  ffn_dtype {dt}: rel diff vs fp32 encoder {rel(f[0], f32[0]):.2e} / {rel(f[1], f32[1]):.2e}")
        TransformerLayer.ffn_dtype = torch.float16

        # CUDA graph of the default path
        static_in = images.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                m.get_img_feat(static_in)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(graph):
                out = m.get_img_feat(static_in)
            graph.replay()
            torch.cuda.synchronize()
            print(f"  graph: rel diff {rel(out[0], ref[0]):.2e} / {rel(out[1], ref[1]):.2e}")
            timeit("encoder CUDA graph replay", graph.replay, args.reps)
        except Exception as e:  # noqa: BLE001
            print("graph capture failed:", repr(e))

        # channels_last convolutions
        if hasattr(enc, "backbone_memory_format"):
            enc.backbone_memory_format = torch.channels_last
            f = m.get_img_feat(images)
            print(f"  channels_last: rel diff {rel(f[0], ref[0]):.2e} / {rel(f[1], ref[1]):.2e}")
            timeit("encoder channels_last", lambda: m.get_img_feat(images), args.reps)
            x = enc.normalize_images(images).reshape(3, 3, H, W)
            timeit("  backbone channels_last", lambda: enc.backbone(x.contiguous(memory_format=torch.channels_last)), args.reps)
            enc.backbone_memory_format = torch.contiguous_format
        x = enc.normalize_images(images).reshape(3, 3, H, W)
        timeit("  backbone default", lambda: enc.backbone(x), args.reps)
        base = enc.backbone(x)
        f0 = torch.stack([base[0], base[0], base[1]])
        f1 = torch.stack([base[1], base[2], base[2]])
        from matchnerf_b200.gmflow import _matmul_precision
        with _matmul_precision("tf32"):
            timeit("  transformer (tf32)", lambda: enc.transformer(f0, f1, 2), args.reps)
        t0, t1 = enc.transformer(f0, f1, 2)
        up_in = torch.cat([t0, t1], 0)
        timeit("  upsampler default", lambda: enc.featup_net(up_in), args.reps)
        if hasattr(enc, "upsampler_memory_format"):
            up_cl = up_in.contiguous(memory_format=torch.channels_last)
            timeit("  upsampler channels_last", lambda: enc.featup_net(up_cl), args.reps)
        timeit("  regroup + pack (get_img_feat tail): full - parts", lambda: None, 1)


if __name__ == "__main__":
    main()
