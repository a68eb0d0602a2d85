#!/usr/bin/env python
"""Relative RMS of the encoder's feature maps against an all-fp32 run of the same modules, per precision setting."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import make_opts
from matchnerf_b200.matchnerf import MatchNeRF
from matchnerf_b200.gmflow import CNNEncoder, TransformerLayer, GMFlow
from oracle import synth
dev = torch.device("cuda", 0)
m = MatchNeRF(make_opts(64, str(dev))).eval()
m.feat_enc.load_state_dict(synth.synthetic_encoder(1)); m.to(dev)
m.encoder_cuda_graph = False
def run(imgs, fast, ffn, mm):
    CNNEncoder.fast_dtype, TransformerLayer.ffn_dtype, GMFlow.matmul_precision = fast, ffn, mm
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    if mm == "fp32":
        torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        f = m.get_img_feat(imgs)
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
    return [x.double() for x in f]
for (H, W) in ((64, 96), (512, 640)):
    g = torch.Generator().manual_seed(4)
    imgs = torch.rand(1, 3, 3, H, W, generator=g).to(dev)
    ref = run(imgs, None, None, "fp32")
    for name, cfg in (("tf32 + fp16 FFN (round-1 default)", (None, torch.float16, "tf32")), ("+ fp16 backbone", (torch.float16, torch.float16, "tf32")),
                      ("fp16 backbone only", (torch.float16, None, "fp32"))):
        f = run(imgs, *cfg)
        print(f"{H}x{W} {name}: " + ", ".join("rel RMS %.2e" % float(((a - b) ** 2).mean().sqrt() / b.std()) for a, b in zip(f, ref)))
