"""GPU (B200): mnf_image_metrics_fwd behind the reference's EvalTools interface (misc/metrics.py) vs the CPU oracle
(oracle/metrics_oracle.py): PSNR within 1e-4 dB, SSIM within 2e-5 (fp32 window arithmetic on both sides)."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("H,W", [(512, 640), (37, 53), (20, 16), (800, 800)])
@pytest.mark.parametrize("masked", [False, True])
def test_eval_tools_vs_oracle(H, W, masked):
    from matchnerf_b200.metrics import EvalTools
    g = np.random.default_rng(H * W + int(masked))
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    gt = np.stack([0.5 + 0.5 * np.sin(xx / 9.0 + c) * np.cos(yy / 7.0) for c in range(3)], -1).astype(np.float32)
    gt = np.clip(gt + 0.05 * g.standard_normal(gt.shape).astype(np.float32), 0, 1)
    pred = np.clip(gt + 0.04 * g.standard_normal(gt.shape).astype(np.float32), -0.1, 1.1).astype(np.float32)   # renders may leave [0, 1]
    mask = (g.random((H, W)) < 0.35) if masked else None
    tools = EvalTools(DEV)
    assert tools.support_metrics[:2] == ["PSNR", "SSIM"]
    want = MO.eval_metrics(pred, gt, mask, return_full=True)
    tools.set_inputs(pred, gt, mask)                                   # numpy inputs: the reference's call (coach.py:430-437)
    got = tools.get_metrics(["PSNR", "SSIM"], return_full=True)
    assert list(got.keys()) == ["PSNR", "PSNR_Full", "SSIM", "SSIM_Full"]
    for k in want:
        tol = 1e-4 if k.startswith("PSNR") else 2e-5
        assert abs(got[k] - want[k]) < tol, (k, got[k], want[k])
    # tensors already on the GPU: nothing but four doubles crosses the bus; identical numbers
    tools.set_inputs(torch.from_numpy(pred).to(DEV), torch.from_numpy(gt).to(DEV), None if mask is None else torch.from_numpy(mask).to(DEV))
    got2 = tools.get_metrics(["PSNR", "SSIM"], return_full=True)
    for k in got:
        assert abs(got[k] - got2[k]) < 1e-9
    # identical images: SSIM 1, PSNR +inf is the reference's behaviour (log 0) -> here a division by zero is not hidden either
    tools.set_inputs(gt, gt, mask)
    assert abs(tools.get_ssim() - 1.0) < 1e-6


def test_image_metrics_rejects_bad_regions(ctx):
    p = torch.rand(16, 16, 3, device=DEV)
    with pytest.raises(RuntimeError):
        ctx.image_metrics(p, p, None, (4, 4, 16, 8))
    with pytest.raises(ValueError):
        ctx.image_metrics(p, torch.rand(16, 15, 3, device=DEV))
