#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).

    python tools/launch_list.py gpurun_out/launches.csv "title" > profiles/rNN_launches_*.md
"""
import csv
import sys


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        rows.append((r[ik], v * scale))
    tot = sum(v for _, v in rows)
    agg = {}
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    print(f"# {title}\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and serialised, compare SHARES.\n")
    print(f"total {tot:.2f} ms over {len(rows)} launches\n")
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k[:90]}` | {n} | {v:.3f} | {100 * v / tot:.1f}% |")


if __name__ == "__main__":
    main()
