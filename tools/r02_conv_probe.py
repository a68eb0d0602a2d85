#!/usr/bin/env python
"""Probe: convolution stacks of the encoder (backbone, up-sampler) as fp32 NCHW (TF32, current) vs fp16 channels_last,
norms excluded (time of the cuDNN calls alone, graph-replayed so that launch gaps do not count)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from matchnerf_b200.gmflow import CNNEncoder, UpSampler

dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True

def graph_time(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def convs_only(enc, x):
    # conv structure of CNNEncoder without the norms (relu instead, to keep magnitudes sane)
    x = F.relu(enc.conv1(x))
    for layer in (enc.layer1, enc.layer2, enc.layer3):
        for blk in layer:
            y = F.relu(blk.conv1(x)); y = F.relu(blk.conv2(y))
            if blk.downsample is not None: x = blk.downsample[0](x)
            x = F.relu(x + y)
    return enc.conv2(x)

with torch.no_grad():
    enc = CNNEncoder().to(dev).eval()
    up = UpSampler().to(dev).eval()
    x = torch.rand(3, 3, 512, 640, device=dev)
    f = torch.randn(6, 128, 64, 80, device=dev)
    print("backbone convs fp32 NCHW (TF32):      %.3f ms" % graph_time(lambda: convs_only(enc, x)))
    xc = x.contiguous(memory_format=torch.channels_last)
    encc = CNNEncoder().to(dev).eval().to(memory_format=torch.channels_last)
    print("backbone convs fp32 channels_last:    %.3f ms" % graph_time(lambda: convs_only(encc, xc)))
    ench = CNNEncoder().to(dev).eval().half().to(memory_format=torch.channels_last)
    xh = xc.half()
    print("backbone convs fp16 channels_last:    %.3f ms" % graph_time(lambda: convs_only(ench, xh)))
    print("backbone full (current path, fused IN kernel): %.3f ms" % graph_time(lambda: enc(x)))
    print("upsampler fp32 NCHW:                  %.3f ms" % graph_time(lambda: up(f)))
    fc = f.contiguous(memory_format=torch.channels_last)
    upc = UpSampler().to(dev).eval().to(memory_format=torch.channels_last)
    print("upsampler fp32 channels_last (current): %.3f ms" % graph_time(lambda: upc(fc)))
    uph = UpSampler().to(dev).eval().half().to(memory_format=torch.channels_last)
    fh = fc.half()
    print("upsampler fp16 channels_last:         %.3f ms" % graph_time(lambda: uph(fh)))
    # elementwise cost reference: one IN-like pass over the biggest activation (3 x 64 x 256 x 320) fp32 vs fp16
    a32 = torch.randn(3, 64, 256, 320, device=dev); a16 = a32.half()
    print("relu pass 3x64x256x320 fp32: %.3f ms, fp16: %.3f ms" % (graph_time(lambda: F.relu(a32)), graph_time(lambda: F.relu(a16))))
