"""Camera paths for video rendering (host-side numpy; nothing here is per-ray).

Behavioural mirror of the two path generators MatchNeRF.get_video_rendering_path uses
(misc/camera.py:382-412 ``get_interpolate_render_path``, :416-468 ``get_spiral_render_path``), restated from scratch:

* interpolate: visit the source cameras in a loop v0 -> v1 -> ... -> v0; on each of the N legs take ``n_frames // 3``
  poses whose xyz-Euler angles (degrees, unwrapped against the FIRST camera's angles) and positions are linear blends
  of the leg's end points (weight 1 -> 0, end point excluded);
* spiral: the LLFF fly-through -- average pose of all scene cameras, focus depth from near / far
  (1 / (0.25 / near + 0.75 / far)), radii = 70th percentile of the camera offsets x ``rads_scale``, two turns.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation


def interpolate_path(c2ws: np.ndarray, n_frames: int = 30) -> np.ndarray:
    """c2ws [V, 3|4, 4] camera-to-world -> [V * (n_frames // 3), 4, 4] (misc/camera.py:382-412)."""
    c2ws = np.asarray(c2ws, dtype=np.float64)
    V = c2ws.shape[0]
    per_leg = n_frames // 3
    w = np.linspace(1.0, 0.0, per_leg, endpoint=False)[:, None]
    angles = np.stack([Rotation.from_matrix(c2ws[i, :3, :3]).as_euler("xyz", degrees=True) for i in range(V)])
    # unwrap every later camera against the first one (the reference adds 360 where the difference exceeds 180 degrees)
    jump = np.abs(angles[1:] - angles[0]) > 180.0
    angles[1:][jump] += 360.0
    centres = c2ws[:, :3, 3]
    nxt = np.roll(np.arange(V), -1)                              # leg i runs from camera i to camera i + 1 (last: back to 0)
    ang = np.concatenate([w * angles[i] + (1.0 - w) * angles[nxt[i]] for i in range(V)])
    pos = np.concatenate([w * centres[i] + (1.0 - w) * centres[nxt[i]] for i in range(V)])
    out = np.tile(np.eye(4), (ang.shape[0], 1, 1))
    out[:, :3, :3] = Rotation.from_euler("xyz", ang, degrees=True).as_matrix()
    out[:, :3, 3] = pos
    return out


def _unit(x):
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def _look_at(z, up, pos):
    """4x4 camera-to-world with viewing axis z, approximate up vector and centre (misc/camera.py:447-455)."""
    z = _unit(z)
    x = _unit(np.cross(up, z))
    y = _unit(np.cross(z, x))
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, pos
    return m


def spiral_path(c2ws_all: np.ndarray, near_far, rads_scale: float = 0.5, n_frames: int = 120, n_rots: int = 2,
                zrate: float = 0.5) -> np.ndarray:
    """All scene cameras [M, 3|4, 4] + (near, far) -> [n_frames, 4, 4] spiral around the average pose
    (misc/camera.py:416-468)."""
    c2ws_all = np.asarray(c2ws_all, dtype=np.float64)
    centre = c2ws_all[:, :3, 3].mean(0)
    up_sum = c2ws_all[:, :3, 1].sum(0)
    avg = _look_at(c2ws_all[:, :3, 2].sum(0), up_sum, centre)
    up = _unit(up_sum)
    near, far = float(near_far[0]), float(near_far[1])
    focal = 1.0 / (0.25 / near + 0.75 / far)
    rads = np.percentile(np.abs(c2ws_all[:, :3, 3] - avg[:3, 3][None]), 70, axis=0) * rads_scale
    rads = np.concatenate([rads, [1.0]])
    poses = []
    for theta in np.linspace(0.0, 2.0 * np.pi * n_rots, n_frames + 1)[:-1]:
        c = avg[:3, :4] @ (np.array([np.cos(theta), -np.sin(theta), -np.sin(theta * zrate), 1.0]) * rads)
        target = avg[:3, :4] @ np.array([0.0, 0.0, -focal, 1.0])
        poses.append(_look_at(c - target, up, c))
    return np.stack(poses)
