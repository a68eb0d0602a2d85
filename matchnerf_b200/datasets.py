"""Dataset loader for forward-facing captures in the LLFF / COLMAP layout (``<root>/<scene>/images/*`` + ``poses_bounds.npy``), with
the reference's sample schema -- the ``datas_dict['colmap']`` entry of datasets/__init__.py (datasets/colmap.py:13-173 on top of
datasets/llff.py:104-159), i.e. what ``test.py --yaml=demo_own`` / ``test_video_own`` read.  Host-side I/O: numpy + PIL, cameras of a
scene computed for all views at once.

A sample (one target view + its ``n_views`` source views, target LAST) is the batch ``MatchNeRF.forward`` consumes:
    images [V+1, 3, H, W] float32 in [0, 1]   extrinsics [V+1, 4, 4] world->camera   intrinsics [V+1, 3, 3]   near_fars [V+1, 2]
    view_ids [V+1]   scene   img_wh   c2ws_all [N_train, 4, 4]
The other loaders of the reference (DTU, Blender, LLFF hold-out, IBRNet, TnT) read datasets that are not available offline; they are
not built.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

IMAGE_EXTENSIONS = (".jpg", ".JPG", ".jpeg", ".JPEG", ".png", ".PNG", ".ppm", ".PPM", ".bmp", ".BMP", ".tif", ".TIF", ".tiff", ".TIFF")
# poses_bounds.npy stores camera-to-world columns as [down, right, back]; the model's convention is OpenCV [right, down, forward]
_FLIP_YZ = np.diag([1.0, -1.0, -1.0, 1.0])
DEPTH_SCALE = 0.47058824          # datasets/colmap.py:104: the nearest bound lands a little above 1 / 0.47 ~ 2.1 after rescaling


def list_images(folder: str) -> List[str]:
    """misc/utils.py:265-275: image file names of a folder, sorted."""
    return sorted(f for f in os.listdir(folder) if f.endswith(IMAGE_EXTENSIONS))


def split_views(cam_positions: np.ndarray, n_select: int = 20, n_interval: int = 6) -> Tuple[np.ndarray, np.ndarray]:
    """(train_views, test_views) of a scene (datasets/colmap.py:13-46): the ``n_select`` cameras closest (L1) to the mean camera
    position, every ``n_interval``-th of them held out for testing; scenes of at most three images test on image 0 from (2, 1, 0)."""
    n = cam_positions.shape[0]
    if n <= 3:
        return np.array([2, 1, 0]), np.array([0])
    n_select, n_interval = min(n, int(n_select)), min(n, int(n_interval))
    order = np.argsort(np.abs(cam_positions - cam_positions.mean(0, keepdims=True)).sum(-1))[:n_select]
    return np.delete(order, range(0, n_select, n_interval)), order[::n_interval]


class PosesBoundsScene:
    """Cameras of one scene from ``poses_bounds.npy`` ([N, 17]: a 3 x 5 block [R | t | (h, w, focal)] + near / far per image),
    datasets/colmap.py:89-131: axes re-ordered to OpenCV, translation and bounds divided by ``DEPTH_SCALE * min(near)``, intrinsics
    rescaled to the requested image size, world-to-camera as the float32 inverse."""

    def __init__(self, scene_dir: str, img_wh: Sequence[int]):
        pb = np.load(os.path.join(scene_dir, "poses_bounds.npy"))
        blocks = pb[:, :15].reshape(-1, 3, 5)
        self.images = list_images(os.path.join(scene_dir, "images"))
        c2w = np.concatenate([blocks[..., 1:2], -blocks[..., :1], blocks[..., 2:4]], -1) @ _FLIP_YZ       # [N, 3, 4]
        bounds = pb[:, -2:].copy()
        scale = bounds.min() * DEPTH_SCALE
        c2w[..., 3] /= scale
        self.near_fars = bounds / scale
        n = c2w.shape[0]
        self.c2w = np.tile(np.eye(4), (n, 1, 1))
        self.c2w[:, :3] = c2w
        self.w2c = np.stack([np.linalg.inv(m.astype(np.float32)) for m in self.c2w])
        w, h = img_wh
        raw_h, raw_w, focal = blocks[:, 0, 4], blocks[:, 1, 4], blocks[:, 2, 4]
        self.K = np.zeros((n, 3, 3))
        self.K[:, 0, 0], self.K[:, 1, 1] = focal * w / raw_w, focal * h / raw_h
        self.K[:, 0, 2], self.K[:, 1, 2], self.K[:, 2, 2] = w / 2, h / 2, 1.0
        # the split is taken on the poses BEFORE the OpenCV flip and the rescaling (gen_pairs, colmap.py:36-40)
        self.train_views, self.test_views = split_views(blocks[..., 3])


class MVSDatasetCOLMAP(torch.utils.data.Dataset):
    """``datas_dict['colmap']`` (datasets/colmap.py:49-173).  ``test_views_method``: 'nearest' = source views sorted by L1 camera
    distance to the target, 'fixed' = the training views in split order and only the first test view (video rendering).
    ``nf_mode``: 'avg' / 'minmax' = one near / far pair for all views of a sample (colmap.py:155-164)."""

    def __init__(self, root_dir, split, n_views=3, img_wh=None, downSample=1.0, max_len=-1, scene_list=None,
                 test_views_method="nearest", nf_mode="avg", **kwargs):
        if split != "test":
            raise AssertionError('Only support "test" split for colmap dataset!')
        if img_wh is None:
            raise ValueError("img_wh = (width, height) is required")
        if nf_mode not in ("avg", "minmax"):
            raise Exception(f"Unknown near far mode {nf_mode}")
        if test_views_method not in ("nearest", "fixed"):
            raise Exception("Unknown evaluate method [%s]" % test_views_method)
        self.root_dir, self.split, self.n_views, self.max_len, self.nf_mode = root_dir, split, int(n_views), int(max_len), nf_mode
        self.img_wh = [int(v) for v in img_wh]
        if scene_list is None:
            scene_list = sorted(x for x in os.listdir(root_dir) if os.path.isdir(os.path.join(root_dir, x)))
        self.scenes: Dict[str, PosesBoundsScene] = {}
        self.metas = []                                    # (scene, target view, source views in order, all training views)
        for name in scene_list:
            sc = PosesBoundsScene(os.path.join(root_dir, name), self.img_wh)
            self.scenes[name] = sc
            tests = sc.test_views[:1] if test_views_method == "fixed" else sc.test_views
            for tgt in tests:
                if test_views_method == "nearest":
                    d = np.abs(sc.c2w[sc.train_views, :3, 3] - sc.c2w[tgt, :3, 3]).sum(-1)
                    src = [sc.train_views[i] for i in np.argsort(d)]
                else:
                    src = sc.train_views
                self.metas.append((name, tgt, src, sc.train_views))

    def get_name(self):
        return "colmap"

    def __len__(self):
        return len(self.metas) if self.max_len <= 0 else self.max_len

    def _load_image(self, scene: str, view: int) -> torch.Tensor:
        from PIL import Image
        path = os.path.join(self.root_dir, scene, "images", self.scenes[scene].images[view])
        img = Image.open(path).resize(tuple(self.img_wh), Image.LANCZOS)
        a = np.array(img.convert("RGB") if img.mode not in ("RGB", "L") else img, dtype=np.uint8)
        if a.ndim == 2:
            a = a[..., None]
        return torch.from_numpy(np.ascontiguousarray(a)).permute(2, 0, 1).float().div(255.0)      # ToTensor: [C, H, W] in [0, 1]

    def __getitem__(self, idx):
        scene, tgt, src, train_views = self.metas[idx]
        sc = self.scenes[scene]
        ids = [int(src[i]) for i in range(self.n_views)] + [int(tgt)]
        nf = sc.near_fars[ids]
        if self.nf_mode == "minmax":
            nf_all = np.array([nf.min() * 0.8, nf.max() * 1.2])
        else:
            nf_all = nf.mean(0)
        return {
            "images": torch.stack([self._load_image(scene, v) for v in ids]).float(),
            "extrinsics": sc.w2c[ids].astype(np.float32),
            "intrinsics": sc.K[ids].astype(np.float32),
            "view_ids": np.array(ids),
            "scene": scene,
            "img_wh": np.array(self.img_wh).astype("int"),
            "near_fars": np.repeat(nf_all[None], len(ids), 0).astype(np.float32),
            "c2ws_all": sc.c2w[np.asarray(train_views)].astype(np.float32),
        }


datas_dict = {"colmap": MVSDatasetCOLMAP}
