"""CPU: this repo's dataset loaders (matchnerf_b200/datasets.py) against the UNMODIFIED reference loaders (datasets/*.py) -- every
field of every compared sample identical, bit for bit: the COLMAP-layout loader on the reference's shipped demo scene and on a
synthetic 14-view scene; the LLFF, Blender and DTU loaders on synthetic dataset trees in the original on-disk formats, with a
synthetic ``configs/pairs.th`` / ``configs/dtu_meta`` both sides read.  The reference's ``torch.load('configs/pairs.th')`` fails on
PyTorch >= 2.6 (weights_only default); for the comparison it is called with the pre-2.6 behaviour.
Needs the reference tree (dev container: /root/reference); skipped without it."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import reference_shim as RS

pytestmark = pytest.mark.skipif(RS.reference_root() is None, reason="reference tree not present")


def _same(a, b):
    assert set(a.keys()) == set(b.keys()), (sorted(a.keys()), sorted(b.keys()))
    for k in a:
        x, y = a[k], b[k]
        if torch.is_tensor(x):
            assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y), k
        elif isinstance(x, np.ndarray):
            assert x.shape == y.shape and x.dtype == y.dtype and np.array_equal(x, y), (k, x, y)
        else:
            assert x == y, k


@pytest.fixture
def legacy_torch_load(monkeypatch):
    real = torch.load
    monkeypatch.setattr(torch, "load", lambda *a, **k: real(*a, **{"weights_only": False, **k}))


def _both(name, root, **kw):
    RS.install_shim()
    ref_ds = importlib.import_module("datasets").datas_dict[name](root, kw.pop("split", "test"), **kw)
    from matchnerf_b200.datasets import datas_dict
    ours = datas_dict[name](root, ref_ds.split, **kw)
    assert len(ours) == len(ref_ds) and ours.get_name() == ref_ds.get_name() == name
    return ours, ref_ds


def _write_images(folder, names, size_hw, g, mode="RGB"):
    from PIL import Image
    folder.mkdir(parents=True, exist_ok=True)
    ch = {"RGB": 3, "RGBA": 4}[mode]
    for nm in names:
        Image.fromarray((g.random((*size_hw, ch)) * 255).astype(np.uint8), mode).save(folder / nm)


def _poses_bounds(n, g):
    pb = np.zeros((n, 17))
    for i in range(n):
        ang = (i - n / 2) * 0.05
        R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        t = np.array([np.sin(ang) * 4, 0.1 * g.standard_normal(), 0.2 * g.standard_normal()])
        pb[i, :15] = np.concatenate([R, t[:, None], np.array([[48.0], [64.0], [55.0]])], 1).ravel()
        pb[i, 15:] = [2.0 + 0.3 * g.random(), 9.0 + g.random()]
    return pb


# ------------------------------------------------------------------------------------------------------------------ COLMAP layout
@pytest.mark.parametrize("method,nf_mode", [("fixed", "minmax"), ("nearest", "avg")])
def test_colmap_demo_scene_matches_reference_loader(method, nf_mode):
    root = os.path.join(RS.reference_root(), "docs/demo_data")
    ours, ref_ds = _both("colmap", root, n_views=3, img_wh=[256, 160], scene_list=["printer"], test_views_method=method, nf_mode=nf_mode)
    assert len(ours) == 1
    _same(ours[0], ref_ds[0])
    assert ours[0]["images"].shape == (4, 3, 160, 256)


def test_colmap_synthetic_scene_matches_reference_loader(tmp_path):
    """14 cameras on a jittered arc (the > 3 image branch: nearest-to-centre selection, every 6th held out) and a second 3-image scene."""
    g = np.random.default_rng(3)
    for scene, n in (("arc", 14), ("tiny", 3)):
        _write_images(tmp_path / scene / "images", [f"img_{i:03d}.png" for i in range(n)], (48, 64), g)
        np.save(tmp_path / scene / "poses_bounds.npy", _poses_bounds(n, g))
    for method, nf_mode in (("nearest", "avg"), ("nearest", "minmax"), ("fixed", "avg")):
        ours, ref_ds = _both("colmap", str(tmp_path), n_views=3, img_wh=[32, 24], test_views_method=method, nf_mode=nf_mode)
        assert len(ours) >= 2
        for i in range(len(ours)):
            _same(ours[i], ref_ds[i])


# ------------------------------------------------------------------------------------------------------------------ LLFF
def test_llff_synthetic_scenes_match_reference_loader(tmp_path, monkeypatch, legacy_torch_load):
    g = np.random.default_rng(5)
    data = tmp_path / "nerf_llff_data"
    pairs = {}
    for scene, n in (("fern", 20), ("room", 17)):
        _write_images(data / scene / "images", [f"IMG_{i:04d}.JPG".replace("JPG", "png") for i in range(n)], (48, 64), g)
        np.save(data / scene / "poses_bounds.npy", _poses_bounds(n, g))
        perm = g.permutation(n)
        pairs[f"{scene}_train"], pairs[f"{scene}_val"], pairs[f"{scene}_test"] = perm[:12], perm[12:15], perm[12:15]
    (tmp_path / "configs").mkdir()
    torch.save(pairs, tmp_path / "configs" / "pairs.th")
    monkeypatch.chdir(tmp_path)                                   # both loaders look for configs/pairs.th under the working directory
    for kw in (dict(test_views_method="nearest", eval_mode="mvsnerf"), dict(test_views_method="fixed", eval_mode="mvsnerf"),
               dict(test_views_method="nearest", eval_mode="gpnr")):
        ours, ref_ds = _both("llff", str(data), n_views=3, img_wh=[32, 24], **kw)
        assert len(ours) >= 4
        for i in range(len(ours)):
            _same(ours[i], ref_ds[i])


# ------------------------------------------------------------------------------------------------------------------ Blender
def test_blender_synthetic_scenes_match_reference_loader(tmp_path, monkeypatch, legacy_torch_load):
    g = np.random.default_rng(7)
    data = tmp_path / "nerf_synthetic"
    pairs = {}

    def frames(part, n):
        out = []
        for i in range(n):
            a, b = g.random() * 6.28, 0.3 + 0.5 * g.random()
            z = np.array([np.cos(a) * np.cos(b), np.sin(a) * np.cos(b), np.sin(b)])
            x = np.cross([0, 0, 1.0], z); x /= np.linalg.norm(x)
            m = np.eye(4)
            m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, np.cross(z, x), z, 4.0 * z
            out.append({"file_path": f"./{part}/r_{i}", "rotation": 0.01, "transform_matrix": m.tolist()})
        return out

    for scene in ("lego", "ship"):
        for part, n in (("train", 24), ("test", 5)):
            _write_images(data / scene / part, [f"r_{i}.png" for i in range(n)], (40, 40), g, "RGBA")
            with open(data / scene / f"transforms_{part}.json", "w") as f:
                json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames(part, n)}, f)
        perm = g.permutation(24)
        pairs[f"{scene}_train"], pairs[f"{scene}_val"] = perm[:16], perm[16:20]
    (tmp_path / "configs").mkdir()
    torch.save(pairs, tmp_path / "configs" / "pairs.th")
    monkeypatch.chdir(tmp_path)
    for kw in (dict(test_views_method="nearest", eval_mode="mvsnerf"), dict(test_views_method="fixed", eval_mode="mvsnerf"),
               dict(test_views_method="nearest", eval_mode="gpnr")):
        ours, ref_ds = _both("blender", str(data), n_views=3, img_wh=[32, 32], **kw)
        assert len(ours) >= 8
        for i in range(len(ours)):
            _same(ours[i], ref_ds[i])
    s = ours[0]
    assert s["images"].shape == (4, 3, 32, 32) and float(s["images"].max()) <= 1.0


# ------------------------------------------------------------------------------------------------------------------ IBRNet
def test_ibrnet_synthetic_collection_matches_reference_loader(tmp_path):
    g = np.random.default_rng(13)
    data = tmp_path / "ibrnet_collected"
    for subset, scenes in (("ibrnet_collected_1", ("a1", "b2")), ("real_iconic_noface", ("c3",))):
        for scene in scenes:
            n = int(g.integers(7, 11))
            _write_images(data / subset / scene / "images", [f"{i:03d}.png" for i in range(n)], (36, 48), g)
            np.save(data / subset / scene / "poses_bounds.npy", _poses_bounds(n, g))
    for split in ("val", "train"):
        ours, ref_ds = _both("ibrnet", str(data), split=split, n_views=3, img_wh=[32, 24])
        assert len(ours) == (3 if split == "val" else sum(sc.c2w.shape[0] for sc in ours.scenes.values()))
        for i in range(0, len(ours), 3):
            torch.manual_seed(100 + i)
            a = ours[i]
            torch.manual_seed(100 + i)
            _same(a, ref_ds[i])


# ------------------------------------------------------------------------------------------------------------------ Tanks and Temples
def test_tnt_synthetic_scenes_match_reference_loader(tmp_path, monkeypatch, legacy_torch_load):
    from PIL import Image
    g = np.random.default_rng(17)
    data = tmp_path / "tnt"
    pairs = {}
    for scene, n, size in (("Family", 19, (54, 96)), ("Horse", 17, (60, 80))):
        (data / scene / "cams_1").mkdir(parents=True)
        (data / scene / "images").mkdir(parents=True)
        for v in range(n):
            ang = v * 0.3
            E = np.eye(4)
            E[:3, :3] = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
            E[:3, 3] = [0.01 * g.standard_normal(), 0.002 * g.standard_normal(), 0.01 + 0.001 * g.standard_normal()]
            K = np.array([[1165.7 + g.random(), 0, 962.8], [0, 1166.1 + g.random(), 541.9], [0, 0, 1]])
            txt = "extrinsic\n" + "\n".join(" ".join(repr(float(x)) for x in r) for r in E) + "\n\nintrinsic\n" + \
                  "\n".join(" ".join(repr(float(x)) for x in r) for r in K) + f"\n\n{0.002 + 0.001 * g.random()} 0.00001 192 {0.02 + 0.01 * g.random()}\n"
            (data / scene / "cams_1" / f"{v:08d}_cam.txt").write_text(txt)
            Image.fromarray((g.random((*size, 3)) * 255).astype(np.uint8)).save(data / scene / "images" / f"{v:08d}.jpg", quality=95)
        perm = g.permutation(n)
        pairs[f"TNT_{scene}_train"], pairs[f"TNT_{scene}_val"] = perm[:12], perm[12:15]
    (tmp_path / "configs").mkdir()
    torch.save(pairs, tmp_path / "configs" / "pairs.th")
    monkeypatch.chdir(tmp_path)
    for kw in (dict(test_views_method="nearest", eval_mode="mvsnerf", nf_mode="avg"), dict(test_views_method="fixed", eval_mode="mvsnerf", nf_mode="minmax"),
               dict(test_views_method="nearest", eval_mode="gpnr", nf_mode="avg")):
        ours, ref_ds = _both("tnt", str(data), n_views=3, img_wh=[64, 32], **kw)
        assert len(ours) >= 5
        for i in range(len(ours)):
            _same(ours[i], ref_ds[i])


# ------------------------------------------------------------------------------------------------------------------ DTU
def _write_pfm(path, a):
    with open(path, "wb") as f:
        f.write(f"Pf\n{a.shape[1]} {a.shape[0]}\n-1.000000\n".encode())
        np.flipud(a).astype("<f4").tofile(f)


def test_dtu_synthetic_tree_matches_reference_loader(tmp_path, monkeypatch, legacy_torch_load):
    g = np.random.default_rng(11)
    data = tmp_path / "mvs_training" / "dtu"
    (data / "Cameras" / "train").mkdir(parents=True)
    for v in range(49):
        ang = v * 0.13
        R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        E = np.eye(4)
        E[:3, :3], E[:3, 3] = R, [30.0 * np.sin(ang) + g.standard_normal(), 5.0 * g.standard_normal(), 600.0 + 20 * g.standard_normal()]
        K = np.array([[361.54125 + g.random(), 0, 82.900625], [0, 360.3975 + g.random(), 66.383875], [0, 0, 1]])
        txt = "extrinsic\n" + "\n".join(" ".join(repr(float(x)) for x in r) for r in E) + "\n\nintrinsic\n" + \
              "\n".join(" ".join(repr(float(x)) for x in r) for r in K) + f"\n\n{425.0 + g.random()} {2.5 + 0.01 * g.random()}\n"
        (data / "Cameras" / "train" / f"{v:08d}_cam.txt").write_text(txt)
    scans = ["scan3", "scan7"]
    for scan in scans:
        _write_images(data / "Rectified" / f"{scan}_train", [f"rect_{v + 1:03d}_{l}_r5000.png" for v in range(49) for l in range(7)], (24, 32), g)
    cfg = tmp_path / "configs" / "dtu_meta"
    cfg.mkdir(parents=True)
    (cfg / "train_all.txt").write_text("\n".join(scans) + "\n")
    (cfg / "val_all.txt").write_text("\n".join(scans) + "\n")
    lines = ["49"]
    for v in range(49):
        others = [int(x) for x in g.permutation(49) if x != v][:10]
        lines += [str(v), "10 " + " ".join(f"{o} {1000.0 * g.random():.2f}" for o in others) + " "]
    (cfg / "view_pairs.txt").write_text("\n".join(lines) + "\n")
    perm = g.permutation(49)
    test_views = np.array(sorted(set(perm[16:19].tolist() + [24])))
    torch.save({"dtu_train": perm[:16], "dtu_test": test_views}, tmp_path / "configs" / "pairs.th")
    for scan in scans:
        (data / "Depths" / scan).mkdir(parents=True)
        for v in test_views:
            d = (g.random((1200, 1600)) * 900).astype(np.float32)
            d[g.random((1200, 1600)) < 0.3] = 0.0                     # the object mask of the evaluation: depth == 0
            _write_pfm(data / "Depths" / scan / f"depth_map_{int(v):04d}.pfm", d)
    monkeypatch.chdir(tmp_path)
    for kw in (dict(split="test", test_views_method="nearest"), dict(split="test", test_views_method="fixed", downSample=0.5), dict(split="val")):
        ours, ref_ds = _both("dtu", str(data), n_views=3, img_wh=[640, 512], **kw)
        assert len(ours) >= 2
        for i in range(len(ours)):
            a, b = ours[i], ref_ds[i]
            _same(a, b)
            assert a["depth"].shape == tuple(a["img_wh"][::-1]) and 0.1 < float((a["depth"] == 0).mean()) < 0.5
    ours, ref_ds = _both("dtu", str(data), n_views=3, img_wh=[640, 512], split="train")
    assert len(ours) == 2 * 49 * 7
    for i in (0, 8, 343, len(ours) - 1):
        torch.manual_seed(i)
        a = ours[i]
        torch.manual_seed(i)
        _same(a, ref_ds[i])
        assert "depth" not in a
