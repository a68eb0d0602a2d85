#!/usr/bin/env python
"""Timeline of the v4 split-window attention kernel on CTA (0, 0) (clock64 stamps recorded inside the kernel; debug aid).

    python tools/attn_trace.py [--shift]
Prints, per role (MMA issuer / producer / softmax warp of column half 0 / 1), the mean number of SM cycles between
consecutive protocol events over the key tiles of the window, and the merged timeline of one tile.
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from matchnerf_b200 import capi  # noqa: E402

EV = {1: "issuer: K(t) landed", 2: "issuer: PV(t-1) complete (S free)", 3: "issuer: QK issued + committed", 4: "issuer: P ready",
      5: "issuer: V(t) landed", 6: "issuer: PV issued + committed",
      20: "producer: QK(t-1) complete (K buffer free)", 21: "producer: masks built, K copy issued", 22: "producer: PV(t-1) complete (V buffer free)",
      10: "softmax: S ready (QK complete seen)", 11: "softmax: sweep 1 done", 12: "softmax: row-max barrier passed",
      13: "softmax: sweep 2 done (P stored)", 14: "softmax: O rescaled, stores waited"}
ROLES = ["MMA issuer (warp 9)", "producer (warp 8)", "softmax warp 0 (columns 0-63)", "softmax warp 4 (columns 64-127)"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shift", action="store_true")
    ap.add_argument("--tile", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = capi.get_context(dev)
    lib = capi.load()
    lib.mnf_debug_attn_trace.argtypes = [C.c_void_p]
    lib.mnf_debug_attn_trace.restype = C.c_int32
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(6, 64 * 80, 128, generator=g).to(dev) for _ in range(3))
    for _ in range(3):
        ctx.window_attn(q, k, v, 64, 80, 2, args.shift, impl=2)
    buf = torch.zeros(4 * 256, dtype=torch.int64, device=dev)
    assert lib.mnf_debug_attn_trace(buf.data_ptr()) == 0
    ctx.window_attn(q, k, v, 64, 80, 2, args.shift, impl=2)
    torch.cuda.synchronize()
    assert lib.mnf_debug_attn_trace(None) == 0
    raw = buf.cpu().view(4, 256)
    merged = []
    for r in range(4):
        ev = [(int(x) >> 8, int(x) & 255) for x in raw[r].tolist() if int(x) != 0]
        if not ev:
            print(ROLES[r], ": no events")
            continue
        print(f"\n{ROLES[r]}: {len(ev)} events, {ev[-1][0] - ev[0][0]} cycles from first to last")
        # mean delta to the previous event, grouped by event id
        acc = {}
        for (t0, _), (t1, e1) in zip(ev[:-1], ev[1:]):
            acc.setdefault(e1, []).append(t1 - t0)
        for e, d in sorted(acc.items()):
            d2 = d[1:] if len(d) > 2 else d                      # skip the first tile (pipeline fill)
            print(f"   -> {EV.get(e, e):55s} mean {sum(d2) / len(d2):8.0f} cycles  (n={len(d2)}, min {min(d2)}, max {max(d2)})")
        merged += [(t, r, e) for t, e in ev]
    merged.sort()
    # one tile of the merged timeline: from the issuer's QK commit of tile `--tile` to the next one
    commits = [t for t, r, e in merged if r == 0 and e == 3]
    if len(commits) > args.tile + 1:
        t0, t1 = commits[args.tile], commits[args.tile + 1]
        print(f"\nmerged timeline of key tile {args.tile} (cycles after the issuer committed its QK; tile period {t1 - t0} cycles):")
        for t, r, e in merged:
            if t0 <= t < t1:
                print(f"   {t - t0:7d}  {ROLES[r][:28]:28s} {EV.get(e, e)}")


if __name__ == "__main__":
    main()
