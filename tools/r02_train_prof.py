"""torch.profiler table of the training step (top CUDA kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from matchnerf_b200.matchnerf import MatchNeRF
from matchnerf_b200.utils import AttrDict
from oracle import synth
import bench

dev = torch.device("cuda", 0)
S, n_rays = 128, 1024
o = bench.make_opts(S, str(dev))
o.nerf.rand_rays_train = n_rays
m = MatchNeRF(o).train()
m.feat_enc.load_state_dict(synth.synthetic_encoder(1)); m.nerf_dec.load_state_dict(synth.synthetic_decoder(0)); m.to(dev)
host = bench.synthetic_batch(100)
opt = torch.optim.AdamW(m.parameters(), lr=5e-4, weight_decay=1e-4)
def step():
    b = AttrDict({k: v.to(dev) for k, v in host.items()})
    opt.zero_grad()
    out = m(b, mode="train")
    gt = b["images"][:, -1].reshape(1, 3, -1).permute(0, 2, 1)[:, out["ray_idx"]]
    torch.nn.functional.mse_loss(out["rgb"], gt).backward()
    torch.nn.utils.clip_grad_norm_(m.feat_enc.parameters(), 1.0)
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
