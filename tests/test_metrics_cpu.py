"""CPU: the metrics oracle (oracle/metrics_oracle.py) against closed forms and a direct float64 window evaluation."""
import numpy as np

from oracle import metrics_oracle as MO


def _direct_ssim(x, y, data_range=2.0, win=7):
    H, W = x.shape
    pad = win // 2
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    n = win * win
    acc = []
    for i in range(pad, H - pad):
        for j in range(pad, W - pad):
            a = x[i - pad:i + pad + 1, j - pad:j + pad + 1].astype(np.float64).ravel()
            b = y[i - pad:i + pad + 1, j - pad:j + pad + 1].astype(np.float64).ravel()
            ua, ub = a.mean(), b.mean()
            va, vb, vab = a.var(ddof=1), b.var(ddof=1), ((a - ua) * (b - ub)).sum() / (n - 1)
            acc.append((2 * ua * ub + c1) * (2 * vab + c2) / ((ua ** 2 + ub ** 2 + c1) * (va + vb + c2)))
    return float(np.mean(acc))


def test_ssim_properties_and_direct_windows():
    g = np.random.default_rng(0)
    x = g.random((24, 31, 3), dtype=np.float32)
    y = np.clip(x + 0.1 * g.standard_normal(x.shape).astype(np.float32), 0, 1)
    assert abs(MO.ssim(x, x) - 1.0) < 1e-6
    assert abs(MO.ssim(x, y) - MO.ssim(y, x)) < 1e-7
    want = np.mean([_direct_ssim(x[..., c], y[..., c]) for c in range(3)])
    assert abs(MO.ssim(x, y) - want) < 2e-6, (MO.ssim(x, y), want)
    # constant images a, b: variances 0 -> SSIM = (2ab + C1) / (a^2 + b^2 + C1)
    a, b = np.full((16, 16, 3), 0.25, np.float32), np.full((16, 16, 3), 0.75, np.float32)
    c1 = (0.01 * 2.0) ** 2
    assert abs(MO.ssim(a, b) - (2 * 0.25 * 0.75 + c1) / (0.25 ** 2 + 0.75 ** 2 + c1)) < 1e-6


def test_eval_metrics_mask_and_crop():
    g = np.random.default_rng(1)
    gt = g.random((40, 50, 3), dtype=np.float32)
    pred = np.clip(gt + 0.05 * g.standard_normal(gt.shape).astype(np.float32), 0, 1)
    m = MO.eval_metrics(pred, gt, None, return_full=True)
    crop = (slice(4, -4), slice(5, -5))
    assert abs(m["PSNR"] - MO.psnr(pred[crop], gt[crop])) < 1e-12 and abs(m["PSNR_Full"] - MO.psnr(pred, gt)) < 1e-12
    mse = float(np.mean((pred.astype(np.float64) - gt) ** 2))
    assert abs(m["PSNR_Full"] + 10 * np.log10(mse)) < 1e-9
    mask = g.random((40, 50)) < 0.3
    mm = MO.eval_metrics(pred, gt, mask)
    keep = ~mask
    assert abs(mm["PSNR"] + 10 * np.log10(np.mean((pred[keep].astype(np.float64) - gt[keep]) ** 2))) < 1e-9
    assert 0 < mm["SSIM"] <= 1 and mm["SSIM"] > m["SSIM_Full"] - 0.5
