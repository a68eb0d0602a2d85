"""fp32 CPU restatement of MatchNeRF's per-ray render path (oracle; test infrastructure only).

Every function cites the reference lines it restates (paths relative to the
reference repo root).  The restatement is deliberately written in a different
shape from the reference (flat [N, .] sample lists, channels-last feature maps,
explicit 4-tap bilinear instead of ``grid_sample``, explicit softmax) so that it
is an independent statement of the arithmetic; ``oracle/make_golden.py`` pins it
against the imported reference.

Conventions
-----------
V source views, S samples per ray, R rays, N = R*S samples (ray-major).
``feats``: list over the two scales of ``[V, h, w, 256]`` fp32 (channels-last).
``images``: ``[V, H, W, 3]`` fp32 in [0, 1].
``w2c``: ``[V, 3, 4]``; ``K``: ``[V, 3, 3]``; ``near_far``: ``[V, 2]``.
Decoder parameters: dict keyed like the reference ``nerf_dec`` state_dict.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------
def invert_pose_f64(w2c: Tensor) -> Tensor:
    """[3,4] world->camera  ->  [3,4] camera->world via a float64 4x4 inverse.

    misc/camera.py:231-240 (``cam2world_legacy``: the square pose is inverted in
    double precision and cast back to float32).
    """
    sq = torch.eye(4, dtype=torch.float64)
    sq[:3, :] = w2c.double()
    return torch.linalg.inv(sq)[:3, :].to(torch.float32)


def cast_rays(H: int, W: int, w2c_tgt: Tensor, K_tgt: Tensor, ray_idx: Tensor) -> Tuple[Tensor, Tensor]:
    """Camera centre [3] and un-normalised ray directions [R,3] for pixel ids ``ray_idx``.

    misc/camera.py:255-278 with ``legacy=True``: pixel (x, y) at integer
    coordinates (no +0.5), ``K^-1 [x, y, 1]`` moved to world space with the
    float64-inverted pose; ``ray = p_world - centre``; row-major pixel order.
    """
    c2w = invert_pose_f64(w2c_tgt)
    Kinv = torch.linalg.inv(K_tgt)
    y = torch.div(ray_idx, W, rounding_mode="floor").to(torch.float32)
    x = (ray_idx % W).to(torch.float32)
    pix = torch.stack([x, y, torch.ones_like(x)], dim=-1)          # [R,3]
    cam = pix @ Kinv.T                                              # img2cam, camera.py:221
    hom = torch.cat([cam, torch.ones_like(cam[:, :1])], dim=-1)     # to_hom, camera.py:204
    p_world = hom @ c2w.T
    centre = c2w[:, 3].clone()                                      # (0,0,0,1) @ c2w^T
    return centre, p_world - centre[None]


def sample_depths(near: float, far: float, S: int, jitter: Optional[Tensor] = None) -> Tensor:
    """Depth values [S] (or [R,S] when ``jitter`` [R,S] in [0,1) is given).

    models/matchnerf.py:163-181 with ``legacy_coord``: shift 0, denominator S-1,
    ``t_i = near + (i + u)/(S-1) * (far - near)``; metric parametrisation.
    """
    i = torch.arange(S, dtype=torch.float32)
    u = i if jitter is None else jitter + i
    return u / float(S - 1) * (far - near) + near


def project_ndc(pts: Tensor, w2c: Tensor, K: Tensor, W: int, H: int, near: float, far: float) -> Tensor:
    """World points [N,3] -> (u, v, z) normalised by (W-1, H-1, near..far) in one source view.

    misc/camera.py:351-379 (``get_coord_ref_ndc``; ``lindisp`` off; no z>0 test).
    """
    cam = pts @ w2c[:, :3].T + w2c[:, 3]
    p = cam @ K.T
    u = p[:, 0] / p[:, 2] / float(W - 1)
    v = p[:, 1] / p[:, 2] / float(H - 1)
    z = (p[:, 2] - near) / (far - near)
    return torch.stack([u, v, z], dim=-1)


# ----------------------------------------------------------------------------
# epipolar gather + grouped cosine similarity
# ----------------------------------------------------------------------------
def bilinear_border(fmap: Tensor, gx: Tensor, gy: Tensor) -> Tensor:
    """Sample channels-last ``fmap`` [h,w,C] at normalised grid coords in [-1,1] -> [N,C].

    Semantics of ``F.grid_sample(mode='bilinear', padding_mode='border',
    align_corners=True)`` as called at models/gmflow/utils.py:134 and
    models/matchnerf.py:245: unnormalise ``((g+1)/2)*(size-1)``, clip to the
    border, 4-tap blend (a tap that falls outside has zero weight).
    """
    h, w, _ = fmap.shape
    ix = ((gx + 1.0) / 2.0) * float(w - 1)
    iy = ((gy + 1.0) / 2.0) * float(h - 1)
    ix = ix.clamp(0.0, float(w - 1))
    iy = iy.clamp(0.0, float(h - 1))
    x0 = ix.floor()
    y0 = iy.floor()
    fx = (ix - x0)[:, None]
    fy = (iy - y0)[:, None]
    x0 = x0.long()
    y0 = y0.long()
    x1 = (x0 + 1).clamp(max=w - 1)
    y1 = (y0 + 1).clamp(max=h - 1)
    t00 = fmap[y0, x0]
    t01 = fmap[y0, x1]
    t10 = fmap[y1, x0]
    t11 = fmap[y1, x1]
    return (t00 * (1 - fx) * (1 - fy) + t01 * fx * (1 - fy) + t10 * (1 - fx) * fy + t11 * fx * fy)


def bilinear_border_local(fmap: Tensor, gx: Tensor, gy: Tensor, radius: int, dilation: int) -> Tensor:
    """``sample_features_by_grid`` with ``local_radius > 0`` (models/gmflow/utils.py:136-162): the mean of the bilinear samples
    at the (2r+1)^2 dilated offsets around the un-normalised coordinate.  The reference re-normalises the offset coordinates with
    ``c = (size + (2r+1)*dilation - 1) / 2`` before handing them to ``F.grid_sample(align_corners=True)``, which un-normalises
    with ``(size - 1) / 2`` -- so every sample position is SCALED by (size - 1) / (size + (2r+1)*dilation - 1).  Reproduced as
    is (same operation order), including that shrink towards the origin.
    """
    h, w, _ = fmap.shape
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    ux = gx * cx + cx                                   # utils.py:141-142 (not clipped)
    uy = gy * cy + cy
    k = 2 * radius + 1
    c2x, c2y = (w + k * dilation - 1) / 2.0, (h + k * dilation - 1) / 2.0       # utils.py:153-154
    acc = None
    for oy in range(-radius, radius + 1):               # generate_window_grid: [2r+1 (y), 2r+1 (x), (x, y)]
        for ox in range(-radius, radius + 1):
            nx = ((ux + float(ox * dilation)) - c2x) / c2x                        # utils.py:155
            ny = ((uy + float(oy * dilation)) - c2y) / c2y
            t = bilinear_border(fmap, nx, ny)
            acc = t if acc is None else acc + t
    return acc / float(k * k)                           # adaptive_avg_pool2d over the window (utils.py:160)


def grouped_cosine(a: Tensor, b: Tensor, groups: int, eps: float = 1e-8) -> Tensor:
    """[N,C] x [N,C] -> [N,groups]: cosine similarity inside each contiguous channel group.

    models/matchnerf.py:266-268 (``reshape(B, G, C/G, ...)`` then
    ``CosineSimilarity(dim=2)``, each norm clamped at 1e-8).
    """
    n, c = a.shape
    ag = a.reshape(n, groups, c // groups)
    bg = b.reshape(n, groups, c // groups)
    na = ag.norm(dim=-1).clamp_min(eps)
    nb = bg.norm(dim=-1).clamp_min(eps)
    return (ag * bg).sum(-1) / (na * nb)


def query_cond(pts: Tensor, feats: Sequence[Tensor], images: Tensor, w2c: Tensor, K: Tensor,
               near_far: Tensor, cos_n_group: Sequence[int], local_radius: int = 0, local_dilation: int = 1) -> Tensor:
    """World points [N,3] -> conditioning vector [N, sum(G)+3V+V] = (cos-sim, colours, masks).

    models/matchnerf.py:209-293.  Per view: project, ``grid = uv*2-1``, gather
    both feature scales and the RGB image, visibility mask = strict
    ``-1 < g < 1`` on both axes.  Per scale: each view's 256 channels are two
    halves (one per pair it takes part in); pairs (v0h0,v1h0), (v0h1,v2h0),
    (v1h1,v2h1) [for V=3]; grouped cosine similarity; mean over pairs.
    Output channel order follows cond_nerf.py:59: feat_info, color_info (view
    major), mask_info.
    """
    V, H, W, _ = images.shape
    sampled = [[] for _ in feats]
    colours, masks = [], []
    for v in range(V):
        ndc = project_ndc(pts, w2c[v], K[v], W, H, float(near_far[v, 0]), float(near_far[v, 1]))
        gx = ndc[:, 0] * 2.0 - 1.0
        gy = ndc[:, 1] * 2.0 - 1.0
        for s, fm in enumerate(feats):
            sampled[s].append(bilinear_border(fm[v], gx, gy) if local_radius <= 0 else
                              bilinear_border_local(fm[v], gx, gy, local_radius, local_dilation))
        colours.append(bilinear_border(images[v], gx, gy))
        inside = (gx > -1.0) & (gx < 1.0) & (gy > -1.0) & (gy < 1.0)
        masks.append(inside.to(torch.float32)[:, None])
    # pair list: models/matchnerf.py:261-264
    pair_ids = [(a, b) for a in range(V - 1) for b in range(a, V - 1)]
    sims = []
    for s, per_view in enumerate(sampled):
        c_half = per_view[0].shape[1] // (V - 1)
        halves = [torch.split(x, c_half, dim=1) for x in per_view]
        acc = 0.0
        for (i, j) in pair_ids:
            acc = acc + grouped_cosine(halves[i][j], halves[j + 1][i], cos_n_group[s])
        sims.append(acc / float(len(pair_ids)))
    return torch.cat(sims + colours + masks, dim=-1)


# ----------------------------------------------------------------------------
# conditional MLP + ray transformer
# ----------------------------------------------------------------------------
def posenc_legacy(x: Tensor, L: int) -> Tensor:
    """[N,3] -> [N, 3+6L]: (x, sin(2^k x) k-major then xyz, cos(...)) without pi.

    models/rfdecoder/cond_nerf.py:108-116 and :56-57.
    """
    freq = 2.0 ** torch.arange(L, dtype=torch.float32)
    spec = (x[:, None, :] * freq[None, :, None]).reshape(x.shape[0], -1)   # [N, L*3], k-major
    return torch.cat([x, spec.sin(), spec.cos()], dim=-1)


def _act(name: str, x: Tensor) -> Tensor:
    if name == "ReLU":
        return torch.relu(x)
    if name == "ELU":
        return torch.nn.functional.elu(x)
    raise ValueError(name)


def ray_attention(dec: Dict[str, Tensor], x: Tensor, row_valid: Tensor) -> Tensor:
    """[R,S,16] -> [R,S,16]: 4-head (d=4) self-attention over the samples of a ray.

    models/rfdecoder/ray_transformer.py:49-79 and :14-26.  ``softmax(q/2 . k)``;
    the mask is applied per *query row* (rows with ``row_valid == 0`` get a
    constant -1e9 score -> uniform attention); residual; LayerNorm(eps 1e-6).
    """
    R, S, D = x.shape
    nh, dk = 4, 4
    q = (x @ dec["ray_attention.w_qs.weight"].T).reshape(R, S, nh, dk).permute(0, 2, 1, 3)
    k = (x @ dec["ray_attention.w_ks.weight"].T).reshape(R, S, nh, dk).permute(0, 2, 1, 3)
    v = (x @ dec["ray_attention.w_vs.weight"].T).reshape(R, S, nh, dk).permute(0, 2, 1, 3)
    scores = (q / (dk ** 0.5)) @ k.transpose(-1, -2)                       # [R,nh,S,S]
    scores = torch.where(row_valid[:, None, :, None] == 0, torch.full_like(scores, -1e9), scores)
    scores = scores - scores.max(dim=-1, keepdim=True).values
    p = scores.exp()
    p = p / p.sum(dim=-1, keepdim=True)
    o = (p @ v).permute(0, 2, 1, 3).reshape(R, S, nh * dk)
    o = o @ dec["ray_attention.fc.weight"].T + x
    mu = o.mean(-1, keepdim=True)
    var = ((o - mu) ** 2).mean(-1, keepdim=True)
    o = (o - mu) / torch.sqrt(var + 1e-6)
    return o * dec["ray_attention.layer_norm.weight"] + dec["ray_attention.layer_norm.bias"]


def raytrans_posenc_table(S: int, d_hid: int = 16) -> Tensor:
    """[S,16] sinusoid table added to raw_alpha when ``raytrans_posenc``. cond_nerf.py:118-127."""
    pos = torch.arange(S, dtype=torch.float64)[:, None]
    j = torch.arange(d_hid, dtype=torch.float64)[None, :]
    ang = pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2.0 * torch.floor(j / 2.0) / d_hid)
    tab = torch.where((torch.arange(d_hid) % 2 == 0)[None, :], ang.sin(), ang.cos())
    return tab.to(torch.float32)


def decoder(dec: Dict[str, Tensor], ndc: Tensor, dir_ref: Tensor, cond: Tensor, S: int, *,
            n_views: int = 3, L_3D: int = 10, skips: Sequence[int] = (4,), raytrans_act: str = "ReLU",
            raytrans_posenc: bool = False, density_maskfill: bool = False) -> Tuple[Tensor, Tensor]:
    """(ndc [N,3], dir_ref [R,3], cond [N,22]) -> rgb [N,3], sigma [N].

    models/rfdecoder/cond_nerf.py:52-100 with ``view_dep`` and ``L_view = 0``:
    ``h = relu((W_i h + b_i) * pts_bias(cond))`` for the trunk, skip concat
    ``[enc, h]`` after the listed layers, 16-d alpha head -> ray attention ->
    sigma; colour head on ``[feature_linear(h), dir]``.
    """
    N = ndc.shape[0]
    R = N // S
    enc = posenc_legacy(ndc, L_3D)
    gate = cond @ dec["pts_bias.weight"].T + dec["pts_bias.bias"]
    h = enc
    n_layers = len([k for k in dec if k.startswith("pts_linears.") and k.endswith(".weight")])
    for i in range(n_layers):
        h = torch.relu((h @ dec[f"pts_linears.{i}.weight"].T + dec[f"pts_linears.{i}.bias"]) * gate)
        if i in skips:
            h = torch.cat([enc, h], dim=-1)
    raw = _act(raytrans_act, h @ dec["alpha_linear.0.weight"].T + dec["alpha_linear.0.bias"])
    raw = raw.reshape(R, S, 16)
    if raytrans_posenc:
        raw = raw + raytrans_posenc_table(S)[None]
    n_valid = cond[:, -n_views:].sum(-1).reshape(R, S)
    att = ray_attention(dec, raw, (n_valid > 1).to(torch.float32))
    a1 = _act(raytrans_act, att @ dec["out_alpha_linear.0.weight"].T + dec["out_alpha_linear.0.bias"])
    sigma = torch.relu(a1 @ dec["out_alpha_linear.2.weight"].T + dec["out_alpha_linear.2.bias"])[..., 0]
    if density_maskfill:
        sigma = torch.where(n_valid < 1, torch.zeros_like(sigma), sigma)
    feat = h @ dec["feature_linear.weight"].T + dec["feature_linear.bias"]
    d = dir_ref[:, None, :].expand(R, S, 3).reshape(N, 3)
    hv = torch.relu(torch.cat([feat, d], dim=-1) @ dec["views_linears.0.weight"].T + dec["views_linears.0.bias"])
    rgb = torch.sigmoid(hv @ dec["rgb_linear.weight"].T + dec["rgb_linear.bias"])
    return rgb, sigma.reshape(N)


def composite(sigma: Tensor, rgb: Tensor, depth: Tensor, setbg_opaque: bool = False):
    """sigma [R,S], rgb [R,S,3], depth [R,S] -> rgb [R,3], depth [R,1], opacity [R,1], prob [R,S].

    models/rfdecoder/nerf.py:101-124 with ``wo_render_interval``: alpha = 1 -
    exp(-sigma); T_i = exp(-sum_{j<i} sigma_j); w = T * alpha.
    """
    alpha = 1.0 - torch.exp(-sigma)
    excl = torch.cumsum(sigma, dim=1) - sigma
    w = torch.exp(-excl) * alpha
    out_rgb = (w[..., None] * rgb).sum(1)
    out_depth = (w * depth).sum(1, keepdim=True)
    opacity = w.sum(1, keepdim=True)
    if setbg_opaque:
        out_rgb = out_rgb + (1.0 - opacity)
    return out_rgb, out_depth, opacity, w


# ----------------------------------------------------------------------------
# the per-slice pipeline
# ----------------------------------------------------------------------------
def render_rays(dec: Dict[str, Tensor], feats: Sequence[Tensor], images: Tensor,
                w2c_src: Tensor, K_src: Tensor, nf_src: Tensor,
                w2c_tgt: Tensor, K_tgt: Tensor, nf_tgt: Tensor,
                ray_idx: Tensor, S: int, *, cos_n_group: Sequence[int] = (2, 8),
                jitter: Optional[Tensor] = None, setbg_opaque: bool = False,
                raytrans_act: str = "ReLU", raytrans_posenc: bool = False,
                density_maskfill: bool = False, return_aux: bool = False, local_radius: int = 0, local_dilation: int = 1):
    """models/matchnerf.py:88-143 (``MatchNeRF.render``) for batch size 1.

    Returns rgb [R,3], depth [R,1], opacity [R,1] (and, with ``return_aux``, the
    intermediate cond / ndc / per-sample rgb / sigma used by the kernel tests).
    """
    V, H, W, _ = images.shape
    R = ray_idx.shape[0]
    centre, ray = cast_rays(H, W, w2c_tgt, K_tgt, ray_idx)
    t = sample_depths(float(nf_tgt[0]), float(nf_tgt[1]), S, jitter)
    t = t[None, :].expand(R, S) if t.dim() == 1 else t
    pts = (centre[None, None, :] + ray[:, None, :] * t[..., None]).reshape(R * S, 3)   # camera.py:281-286
    cond = query_cond(pts, feats, images, w2c_src, K_src, nf_src, cos_n_group, local_radius, local_dilation)
    ndc0 = project_ndc(pts, w2c_src[0], K_src[0], W, H, float(nf_src[0, 0]), float(nf_src[0, 1]))
    unit = ray / ray.norm(dim=-1, keepdim=True).clamp_min(1e-12)                        # matchnerf.py:130
    dir_ref = unit @ w2c_src[0, :, :3].T                                               # matchnerf.py:131
    rgb_s, sigma = decoder(dec, ndc0, dir_ref, cond, S, n_views=V, raytrans_act=raytrans_act,
                           raytrans_posenc=raytrans_posenc, density_maskfill=density_maskfill)
    rgb, depth, opacity, prob = composite(sigma.reshape(R, S), rgb_s.reshape(R, S, 3), t, setbg_opaque)
    if return_aux:
        return rgb, depth, opacity, dict(cond=cond, ndc=ndc0, dir_ref=dir_ref, rgb_s=rgb_s, sigma=sigma,
                                         depth_samples=t, prob=prob, pts=pts)
    return rgb, depth, opacity


def to_channels_last(feats_ref: Sequence[Tensor]) -> List[Tensor]:
    """Reference layout [1,V,C,h,w] (matchnerf.py:192-205) -> oracle layout [V,h,w,C]."""
    return [f[0].permute(0, 2, 3, 1).contiguous() for f in feats_ref]
