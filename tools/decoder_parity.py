#!/usr/bin/env python
"""Print the tcgen05 decoder's error against the fp32 oracle (per-sample sigma / rgb and composited outputs) for every
supported S, on the synthetic config-1 scene.  TEST TOOL (uses oracle/).  MNF_LIB_PATH selects a library variant.

    python tools/decoder_parity.py [--rays 512]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from matchnerf_b200 import capi  # noqa: E402
from oracle import synth  # noqa: E402
from helpers import config1_inputs, rms, max_abs  # noqa: E402
import test_gpu_kernels as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=512)
    args = ap.parse_args()
    ctx = capi.get_context(torch.device("cuda", 0))
    feats, imgs, extr, intr, nf, ray_idx = config1_inputs()
    print("library:", capi.LIB_PATH)
    for S in (16, 32, 64, 128):
        for impl in (1, 2):
            got, o = T.run_decoder_case(ctx, synth.synthetic_decoder(0), feats, imgs, extr, intr, nf, ray_idx[:args.rays], S, impl)
            print(f"S={S:3d} impl={impl}: rgb rms {rms(got[0], o[0]):.2e} max {max_abs(got[0], o[0]):.2e} | depth rms "
                  f"{rms(got[1], o[1][:, 0]):.2e} | opacity rms {rms(got[2], o[2][:, 0]):.2e} | sigma rms "
                  f"{rms(got[3][:, 3], o[3]['sigma']):.2e} (mean {float(o[3]['sigma'].mean()):.3f}) | rgb_s rms {rms(got[3][:, :3], o[3]['rgb_s']):.2e}")


if __name__ == "__main__":
    main()
