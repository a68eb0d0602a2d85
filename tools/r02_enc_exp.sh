#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_enc_exp}.log
(timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
 python tools/prof_kernels.py --rays 40960 --which encoder --reps 5 2>&1 | head -4
 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train-step | tail -1 > gpurun_out/${1:-r02_enc_exp}.json
 python - <<PY
import json
d=json.load(open("gpurun_out/${1:-r02_enc_exp}.json"))
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "enc", d["kernels"]["encoder"]["ms_per_call"])
p=d["parity"]
for nm,q in (("S64",p),("S128",p["other_S"])):
    for k in ("vs_cpu_oracle","vs_reference_gpu_fp32"):
        print(nm,k,"rgb_rms %.2e dpsnr %.5f" % (q[k]["rgb_rms"], q[k]["psnr_delta_db"]))
PY
) > $L 2>&1
tail -16 $L | cut -c1-250
