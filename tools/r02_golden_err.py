import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from tests.helpers import load_npz, rms
from tests.test_gpu_model import build_model
from matchnerf_b200.gmflow import CNNEncoder
z = load_npz(os.path.join(ROOT, "tests", "golden"), "encoder_64x96.npz")
scale = float(np.std(z["feat8"]))
for fd in (None, torch.float32, torch.float16):
    CNNEncoder.fast_dtype = fd
    m, opt = build_model(16)
    m.encoder_cuda_graph = False
    with torch.no_grad():
        f8, f4 = m.get_img_feat(torch.from_numpy(z["images"]).to("cuda:0"))
        f8b, _ = m.get_img_feat(torch.from_numpy(z["images"]).to("cuda:0"))
    print(fd, "feat8 rel rms %.3e  feat4 rel rms %.3e  run-to-run equal %s" % (rms(f8[0], z["feat8"]) / scale, rms(f4[0][:, ::8], z["feat4_ch0mod8"]) / scale, torch.equal(f8, f8b)))
