"""One eager encoder forward at DTU size between cudaProfilerStart/Stop, for an ncu launch list:
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/enc_launches.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import synth  # noqa: E402
from bench import make_opts  # noqa: E402
from matchnerf_b200.matchnerf import MatchNeRF  # noqa: E402

dev = torch.device("cuda", 0)
m = MatchNeRF(make_opts(64, str(dev))).eval()
m.feat_enc.load_state_dict(synth.synthetic_encoder(1))
m.to(dev)
m.encoder_cuda_graph = False
images = torch.rand(1, 3, 3, 512, 640, generator=torch.Generator().manual_seed(7)).to(dev)
with torch.no_grad():
    for _ in range(3):
        m.get_img_feat(images)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m.get_img_feat(images)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
