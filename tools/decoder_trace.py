#!/usr/bin/env python
"""Timeline of the tcgen05 decoder kernel on CTA 0 (clock64 stamps recorded inside the kernel; debug aid).

    python tools/decoder_trace.py [--rays 40960] [--samples 64]
Prints, per role (mma / trunk slot / ray group), the mean number of SM cycles between consecutive protocol events.
"""
import argparse
import collections
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from matchnerf_b200 import capi  # noqa: E402
from oracle import synth  # noqa: E402

TRUNK = {0: "tile start", 1: "staged (A operands written)", 2: "gate acc ready", 3: "trunk done, wait heads acc", 4: "heads acc ready",
         5: "heads epilogue done", 6: "hand-off buffer free", 7: "tile end"}
TRUNK.update({30: "staging: cp.async issued", 31: "staging: ray cast", 32: "staging: sample point", 33: "staging: NDC projected", 34: "staging: direction done"})
TRUNK.update({10 + l: f"L{l}: gate/prev epilogue done -> wait acc" for l in range(6)})
TRUNK.update({20 + l: f"L{l}: acc ready" for l in range(6)})
MMA = {}
MMA.update({50 + p: f"ph{p}: weights resident" for p in range(8)})
MMA.update({10 + p: f"ph{p}: slot A operand ready" for p in range(8)})
MMA.update({30 + p: f"ph{p}: MMAs issued + committed" for p in range(8)})
MMA.update({60 + p: f"ph{p}: weight stages released" for p in range(8)})
RAY = {0: "wait hand-off", 1: "hand-off received", 2: "q/k/v projected (phase A)", 3: "keys / values of the ray visible", 4: "attention of an m-tile done", 8: "fc + LN + sigma head of an m-tile done", 5: "ray transformer (mma.sync) done", 6: "sigma read back, buffer released",
       7: "composite + tile end"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=40960)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--quarters", action="store_true", help="also summarise warps 1..3 of every ray group (drift between the warps of a group)")
    ap.add_argument("--timeline", type=int, default=-1, help="also print the merged event timeline of this tile iteration of CTA 0")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = capi.get_context(dev)
    lib = capi.load()
    lib.mnf_debug_decoder_trace.argtypes = [C.c_void_p, C.c_int32]
    lib.mnf_debug_decoder_trace.restype = C.c_int32
    H, W, S = 512, 640, args.samples
    from bench import make_opts
    from matchnerf_b200.matchnerf import MatchNeRF
    feats, imgs, _ = synth.synthetic_scene(H, W, seed=1234)
    extr, intr, nf = synth.synthetic_cameras(H, W)
    m = MatchNeRF(make_opts(S, str(dev))).eval()
    m.feat_enc.load_state_dict(synth.synthetic_encoder(1))
    m.to(dev)
    with torch.no_grad():
        f = m.get_img_feat(imgs.to(dev))
    ctx.load_decoder(synth.synthetic_decoder(0))
    packed = ctx.pack_scene([f[0][0], f[1][0]], imgs[0].to(dev), extr[0, :3], intr[0, :3], nf[0, :3])
    sc = packed.c_scene(extr[0, 3, :3], intr[0, 3], nf[0, 3])
    cfg = capi.DecoderCfg()
    cfg.n_samples, cfg.raytrans_act, cfg.raytrans_posenc, cfg.density_maskfill = S, 0, 0, 0
    first = min(100 * W, H * W - args.rays)
    _, c16 = ctx.gather_cossim(sc, S, first_ray=first, n_rays=args.rays, want_f32=False, want_f16=True)
    ctx.decoder_composite(sc, cfg, cond_f16=c16, first_ray=first, n_rays=args.rays, impl=2)      # warm-up
    torch.cuda.synchronize()
    cap = 20 * 2048
    buf = torch.zeros(cap, dtype=torch.int64, device=dev)
    assert lib.mnf_debug_decoder_trace(buf.data_ptr(), cap) == 0
    ctx.decoder_composite(sc, cfg, cond_f16=c16, first_ray=first, n_rays=args.rays, impl=2)
    torch.cuda.synchronize()
    lib.mnf_debug_decoder_trace(None, 0)
    rec = buf.cpu().numpy().astype("uint64")
    rec = rec[rec != 0]
    ev = [(int(r >> 24), int((r >> 20) & 15), int((r >> 16) & 15), int((r >> 8) & 255), int(r & 255)) for r in rec]
    print(f"{len(ev)} records")
    t0 = min(e[0] for e in ev)
    per = collections.defaultdict(list)
    for clk, role, slot, e, it in ev:
        if role == 0 or (slot >> 2) == 0:
            per[(role, slot & 3)].append((clk - t0, e, it))
        elif args.quarters and role == 2:
            per[(role, (slot & 3) + 10 * (slot >> 2))].append((clk - t0, e, it))   # key 10 q + slot: the other warps of a ray group
    if args.timeline >= 0:
        starts = sorted(c for c, e, it in per[(1, 0)] if e == 0 and it in (args.timeline, args.timeline + 1))
        if len(starts) == 2:
            print(f"\n== merged timeline of iteration {args.timeline} (cycles since slot 0 tile start)")
            allev = sorted((clk - t0, role, slot, e) for clk, role, slot, e, it in ev)
            tag = {0: "MMA", 1: "TRUNK", 2: "RAY"}
            for c, role, slot, e in allev:
                if starts[0] <= c < starts[1]:
                    names = TRUNK if role == 1 else (RAY if role == 2 else MMA)
                    label = names.get(e, e)
                    if role == 0:
                        label = str(label).replace("slot A", "slot %s" % "AB"[slot & 1])
                    print(f"   {c - starts[0]:7d}  {tag[role]:5s} {slot & 3} q{slot >> 2}  {label}")
    for key in sorted(per):
        role, slot = key
        seq = sorted(per[key])
        name = {0: "mma issuer", 1: "trunk slot", 2: "ray group"}[role]
        print(f"\n== {name} {slot}: {len(seq)} events, span {seq[-1][0] - seq[0][0]} cycles")
        # mean delta to the previous event of the same role, keyed by (prev event -> event)
        acc = collections.OrderedDict()
        for (c0, e0, i0), (c1, e1, i1) in zip(seq[:-1], seq[1:]):
            if not (3 <= (i1 if role else 0) or role == 0):
                continue
            acc.setdefault((e0, e1), []).append(c1 - c0)
        names = TRUNK if role == 1 else (RAY if role == 2 else MMA)
        tot = 0.0
        for (e0, e1), v in acc.items():
            if len(v) < 4:
                continue
            mean = sum(v) / len(v)
            tot += mean
            print(f"   {mean:9.0f} cyc  x{len(v):4d}   {e0:3d} -> {e1:3d}   {names.get(e0, e0)}  ->  {names.get(e1, e1)}")
        print(f"   sum of means {tot:.0f} cycles")


if __name__ == "__main__":
    main()
