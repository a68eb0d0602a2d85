"""Time mnf_token_block_fwd at DTU size (6 x 64 x 80 tokens) with CUDA events: tools/time_token_block.py [--rows N]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from matchnerf_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=6 * 64 * 80)
ap.add_argument("--reps", type=int, default=50)
a = ap.parse_args()
ctx = capi.get_context("cuda:0")
dev = ctx.device
g = torch.Generator().manual_seed(0)
mk = lambda *s: torch.randn(*s, generator=g).to(dev)
attn, src = mk(a.rows, 128), mk(a.rows, 128)
for ffn in (False, True):
    ws = [mk(128, 128) / 11.3, 1 + 0.1 * mk(128), 0.1 * mk(128)]
    if ffn:
        ws += [mk(1024, 256) / 16, mk(128, 1024) / 32, 1 + 0.1 * mk(128), 0.1 * mk(128)]
    blob = ctx.token_block_pack(*ws)
    for _ in range(5):
        ctx.token_block(attn, src, blob, ffn)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.reps):
        ctx.token_block(attn, src, blob, ffn)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / a.reps * 1e3
    flop = a.rows * 2 * (128 * 128 + (256 * 1024 + 1024 * 128 if ffn else 0))
    print(f"token_block rows={a.rows} ffn={ffn}: {us:.1f} us per call, {flop / us / 1e6:.1f} TFLOP/s")
