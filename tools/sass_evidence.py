#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-native path: counts of the tcgen05 / TMEM / TMA mnemonics in the in-tree library
(cuobjdump -sass; no GPU needed).  Writes a markdown table to stdout:  python tools/sass_evidence.py > profiles/rNN_sass_evidence.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "matchnerf_b200", "libmatchnerf_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "FFMA2", "REDUX", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("mnf::", "")
            cur = re.sub(r"\(.*", "", name)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            kernels[cur]["_n"] += 1
            for mn in MNEMONICS:
                if m.group(1).startswith(mn):
                    kernels[cur][mn] += 1
    print("# SASS evidence (cuobjdump -sass matchnerf_b200/libmatchnerf_b200.so, sm_100a)\n")
    print("`UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st (tensor memory), `UTMALDG` = cp.async.bulk.tensor (TMA tensor map), "
          "`UBLKCP` = cp.async.bulk (TMA unit, linear), `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier, `HMMA` = mma.sync (legacy tensor path, "
          "used only for the 16-wide ray transformer inside the decoder), `LDGSTS` = cp.async.\n")
    print("| kernel | instructions | " + " | ".join(MNEMONICS) + " |")
    print("|---|---|" + "---|" * len(MNEMONICS))
    for k, c in kernels.items():
        print(f"| `{k}` | {c['_n']} | " + " | ".join(str(c[m]) if c[m] else "" for m in MNEMONICS) + " |")


if __name__ == "__main__":
    sys.exit(main())
