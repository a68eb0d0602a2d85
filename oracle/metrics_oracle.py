"""CPU restatement of the reference's evaluation metrics (TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's cpu_baseline leg may
import this).

misc/metrics.py:19-46 computes PSNR with numpy and SSIM with ``skimage.metrics.structural_similarity(pred, gt, channel_axis=-1)``.
scikit-image (pinned to 0.19.2 by the reference's requirements.txt) is a third-party dependency that is ABSENT from this image and
from /root/reference, so its published algorithm (skimage/metrics/_structural_similarity.py of 0.19.2; Wang et al. 2004 with a
uniform window) is restated here on top of ``scipy.ndimage.uniform_filter`` -- the very routine skimage calls.  PARITY UNPINNED for
SSIM against skimage itself (no golden vector of it can be produced here); pinned properties: SSIM(x, x) = 1, symmetry, the
closed-form value for constant images, and a direct O(49) window evaluation in float64.  PSNR is the formula of misc/metrics.py:35-41.
LPIPS (misc/metrics.py:47-52) needs the `lpips` package and its VGG weights: not available offline, not restated.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import uniform_filter


def ssim_channel(x: np.ndarray, y: np.ndarray, data_range: float = 2.0, win: int = 7, k1: float = 0.01, k2: float = 0.03) -> float:
    """One channel, skimage 0.19.2 defaults (gaussian_weights=False, use_sample_covariance=True): mean of the SSIM map with the
    (win - 1) // 2 border cropped.  x, y: [H, W] float32 (skimage keeps float32 inputs in float32)."""
    npix = win * win
    cov_norm = npix / (npix - 1.0)
    ux, uy = uniform_filter(x, size=win), uniform_filter(y, size=win)
    uxx, uyy, uxy = uniform_filter(x * x, size=win), uniform_filter(y * y, size=win), uniform_filter(x * y, size=win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return float(s[pad:-pad, pad:-pad].mean(dtype=np.float64))


def ssim(pred: np.ndarray, gt: np.ndarray, data_range: float = 2.0) -> float:
    """``structural_similarity(pred, gt, channel_axis=-1)``: the mean over the channels of the per-channel SSIM.  data_range is
    NOT passed by the reference, so skimage 0.19 takes the dtype range of float images, (-1, 1) -> 2."""
    return float(np.mean([ssim_channel(np.ascontiguousarray(pred[..., c], np.float32), np.ascontiguousarray(gt[..., c], np.float32),
                                       data_range) for c in range(pred.shape[-1])]))


def psnr(pred: np.ndarray, gt: np.ndarray, keep: np.ndarray | None = None) -> float:
    """misc/metrics.py:35-41: -10 log10(mean squared error) over all (or the kept) pixels."""
    d = (pred - gt) if keep is None else (pred[keep] - gt[keep])
    return float(-10.0 * np.log(np.mean(d.astype(np.float64) ** 2)) / np.log(10.0))


def eval_metrics(pred: np.ndarray, gt: np.ndarray, mask: np.ndarray | None = None, return_full: bool = False) -> dict:
    """EvalTools.set_inputs + get_metrics(['PSNR', 'SSIM']) (misc/metrics.py:19-33, :54-65).  mask: True = masked OUT."""
    out = {}
    if mask is not None:
        pp, pg = pred.copy(), gt.copy()
        pp[mask] = 0.0
        pg[mask] = 0.0
        out["PSNR"] = psnr(pp, pg, ~mask)
        out["SSIM"] = ssim(pp, pg)
    else:
        hc, wc = np.array(pred.shape[:2]) // 10
        pp, pg = pred[hc:-hc, wc:-wc], gt[hc:-hc, wc:-wc]
        out["PSNR"] = psnr(pp, pg)
        out["SSIM"] = ssim(pp, pg)
    if return_full:
        out["PSNR_Full"] = psnr(pred, gt)
        out["SSIM_Full"] = ssim(pred, gt)
    return out
