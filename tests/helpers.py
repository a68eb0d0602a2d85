"""Shared test helpers: golden loading, oracle adapters, comparison metrics."""
import os

import numpy as np
import torch

from oracle import render_oracle as RO
from oracle import synth


def rms(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float(((a - b) ** 2).mean().sqrt())


def max_abs(a, b):
    return float((torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu()).abs().max())


def frac_above(a, b, tol):
    d = (torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu()).abs()
    return float((d > tol).double().mean())


def psnr(a, b):
    """misc/metrics.py:35-41 formula: -10 log10(mean((a-b)^2))."""
    mse = float(((torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu()) ** 2).mean())
    return -10.0 * np.log10(max(mse, 1e-20))


def load_npz(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def dec_from_npz(z):
    return {k[4:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("dec.")}


def half_round(t):
    return t.to(torch.float16).to(torch.float32)


def config1_inputs(H=512, W=640):
    feats, imgs, g = synth.synthetic_scene(H, W, seed=1234)
    extr, intr, nf = synth.synthetic_cameras(H, W)
    ray_idx = torch.randperm(H * W, generator=g)[:1024]
    return feats, imgs, extr, intr, nf, ray_idx


def oracle_render(dec, feats, imgs, extr, intr, nf, ray_idx, S, quantize_feats=False, **kw):
    """feats: reference layout [1,V,C,h,w] list; imgs [1,V,3,H,W]."""
    f = [half_round(x) for x in feats] if quantize_feats else feats
    return RO.render_rays(dec, RO.to_channels_last(f), imgs[0].permute(0, 2, 3, 1).contiguous(),
                          extr[0, :3, :3], intr[0, :3], nf[0, :3], extr[0, 3, :3], intr[0, 3], nf[0, 3], ray_idx, S, **kw)
