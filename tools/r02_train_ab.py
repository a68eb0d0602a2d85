"""A/B of the training step (bench.train_step_times) with the SDPA attention on / off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import matchnerf_b200.train_path as TP
dev = torch.device("cuda", 0)
for flag in (True, False, True):
    TP.USE_SDPA = flag
    r = bench.train_step_times(dev, steps=5, warmup=2)
    print("USE_SDPA", flag, "ours %.2f ms" % r["ours_ms_per_step"], "loss %.5f" % r["ours_last_loss"], "reference %.2f ms" % r.get("reference_gpu_ms_per_step", float("nan")))
