"""CPU: this repo's LLFF / COLMAP-layout loader (matchnerf_b200/datasets.py) against the UNMODIFIED reference loader
(datasets/colmap.py) on the reference's shipped demo scene and on a synthetic 14-view scene: every field of every sample identical.
Needs the reference tree (dev container: /root/reference); skipped without it."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_shim as RS

pytestmark = pytest.mark.skipif(RS.reference_root() is None, reason="reference tree not present")


def _same(a, b):
    assert set(a.keys()) == set(b.keys())
    for k in a:
        x, y = a[k], b[k]
        if torch.is_tensor(x):
            assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y), k
        elif isinstance(x, np.ndarray):
            assert x.shape == y.shape and x.dtype == y.dtype and np.array_equal(x, y), (k, x, y)
        else:
            assert x == y, k


def _both(root, **kw):
    ref = RS.install_shim()
    import importlib
    ref_ds = importlib.import_module("datasets").datas_dict["colmap"](root, "test", **kw)
    from matchnerf_b200.datasets import datas_dict
    ours = datas_dict["colmap"](root, "test", **kw)
    assert len(ours) == len(ref_ds) and ours.get_name() == ref_ds.get_name()
    return ours, ref_ds


@pytest.mark.parametrize("method,nf_mode", [("fixed", "minmax"), ("nearest", "avg")])
def test_demo_scene_matches_reference_loader(method, nf_mode):
    root = os.path.join(RS.reference_root(), "docs/demo_data")
    ours, ref_ds = _both(root, n_views=3, img_wh=[256, 160], scene_list=["printer"], test_views_method=method, nf_mode=nf_mode)
    assert len(ours) == 1
    _same(ours[0], ref_ds[0])
    assert ours[0]["images"].shape == (4, 3, 160, 256)


def test_synthetic_scene_matches_reference_loader(tmp_path):
    """14 cameras on a jittered arc (the > 3 image branch: nearest-to-centre selection, every 6th held out) and a second 3-image scene."""
    from PIL import Image
    g = np.random.default_rng(3)
    for scene, n in (("arc", 14), ("tiny", 3)):
        d = tmp_path / scene / "images"
        d.mkdir(parents=True)
        pb = np.zeros((n, 17))
        for i in range(n):
            ang = (i - n / 2) * 0.05
            R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]) @ np.diag([1.0, 1.0, 1.0])
            t = np.array([np.sin(ang) * 4, 0.1 * g.standard_normal(), 0.2 * g.standard_normal()])
            pb[i, :15] = np.concatenate([R, t[:, None], np.array([[48.0], [64.0], [55.0]])], 1).ravel()
            pb[i, 15:] = [2.0 + 0.3 * g.random(), 9.0 + g.random()]
            Image.fromarray((g.random((48, 64, 3)) * 255).astype(np.uint8)).save(d / f"img_{i:03d}.png")
        np.save(tmp_path / scene / "poses_bounds.npy", pb)
    for method, nf_mode in (("nearest", "avg"), ("nearest", "minmax"), ("fixed", "avg")):
        ours, ref_ds = _both(str(tmp_path), n_views=3, img_wh=[32, 24], test_views_method=method, nf_mode=nf_mode)
        assert len(ours) >= 2
        for i in range(len(ours)):
            _same(ours[i], ref_ds[i])
