"""Evaluation metrics with the reference's ``EvalTools`` interface (misc/metrics.py), evaluated on the GPU.

The reference copies the rendered image and the ground truth to the host and calls numpy (PSNR) and scikit-image (SSIM) there; here
both stay on the device and ONE launch of ``mnf_image_metrics_fwd`` per region returns the sums both numbers are made of.  Inputs may
be numpy arrays (the reference's call, coach.py:430-437) or torch tensors already on the GPU (``var['rgb']``, ``var.images[:, -1]``).
LPIPS (misc/metrics.py:47-52) needs the third-party ``lpips`` package with its VGG weights; it is used when importable and reported
as unavailable otherwise (there is no offline copy of the weights).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from . import capi


class EvalTools:
    """misc/metrics.py:10-65.  ``set_inputs(pred_img, gt_img, img_mask=None)`` then ``get_metrics(metrics=None, return_full=False)``;
    images [H, W, 3] in [0, 1], mask [H, W] bool with True = masked out (DTU: depth == 0).  Without a mask the metrics are taken
    on the centre crop to 80 % (metrics.py:29-33)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.ctx = capi.get_context(self.device)
        self.lpips_metric = None
        try:                                               # third-party, not in this image: optional
            import lpips  # type: ignore
            with torch.no_grad():
                self.lpips_metric = lpips.LPIPS(net="vgg").to(self.device)
        except Exception:                                  # package or weights missing
            self.lpips_metric = None
        self.support_metrics = ["PSNR", "SSIM"] + (["LPIPS"] if self.lpips_metric is not None else [])

    def _dev(self, a, dtype=torch.float32):
        t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
        return t.detach().to(self.device, dtype).contiguous()

    def set_inputs(self, pred_img, gt_img, img_mask=None):
        self.full_pred, self.full_gt = self._dev(pred_img), self._dev(gt_img)
        if self.full_pred.dim() != 3 or self.full_pred.shape[-1] != 3 or self.full_pred.shape != self.full_gt.shape:
            raise ValueError("pred_img / gt_img must be [H, W, 3] with equal shapes")
        H, W = self.full_pred.shape[:2]
        self.img_mask = None if img_mask is None else self._dev(img_mask, torch.bool)
        if self.img_mask is None:                           # centre crop to 80 % (metrics.py:29-33)
            hc, wc = H // 10, W // 10
            if hc == 0 or wc == 0:
                raise ValueError("image smaller than 10 pixels: the reference's centre crop is empty")
            self.region = (hc, wc, H - 2 * hc, W - 2 * wc)
        else:
            self.region = (0, 0, H, W)
        self._sums = {}

    def _get_sums(self, full: bool):
        if full not in self._sums:
            s = self.ctx.image_metrics(self.full_pred, self.full_gt, None if full else self.img_mask, None if full else self.region)
            self._sums[full] = s.cpu().tolist()             # the only device -> host traffic: four doubles
        return self._sums[full]

    def get_psnr(self, full: bool = False, **kwargs) -> float:
        sq, n, _, _ = self._get_sums(full)
        if n == 0:
            raise ValueError("no unmasked pixel to evaluate")
        return float("inf") if sq == 0 else -10.0 * math.log(sq / n) / math.log(10.0)     # identical images: numpy's -log(0) = inf

    def get_ssim(self, full: bool = False, **kwargs) -> float:
        _, _, s, n = self._get_sums(full)
        return s / n

    @torch.no_grad()
    def get_lpips(self, full: bool = False, **kwargs) -> float:
        if self.lpips_metric is None:
            raise RuntimeError("LPIPS needs the third-party `lpips` package and its VGG weights (not available offline)")
        p, g = self.full_pred, self.full_gt
        if not full:
            if self.img_mask is not None:
                keep = (~self.img_mask)[..., None].to(p.dtype)
                p, g = p * keep, g * keep
            else:
                y0, x0, h, w = self.region
                p, g = p[y0:y0 + h, x0:x0 + w], g[y0:y0 + h, x0:x0 + w]
        to = lambda t: t[None].permute(0, 3, 1, 2) * 2 - 1.0          # RGB in [-1, 1] (metrics.py:49-50)
        return float(self.lpips_metric(to(p), to(g)).item())

    def get_metrics(self, metrics=None, return_full: bool = False):
        out = OrderedDict()
        for metric in (self.support_metrics if metrics is None else metrics):
            if metric not in ("PSNR", "SSIM", "LPIPS"):
                raise AssertionError("only support metrics: [PSNR,SSIM,LPIPS]")
            fn = getattr(self, f"get_{metric.lower()}")
            out[metric] = fn(full=False)
            if return_full:
                out[f"{metric}_Full"] = fn(full=True)
        return out
