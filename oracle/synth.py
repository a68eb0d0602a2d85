"""Deterministic synthetic parameters and inputs shared by the oracle tests, the golden
generator and the benchmarks (test infrastructure; no reference code involved).

Weights are drawn per parameter from a generator seeded by (seed, parameter name), so the
same state_dict can be loaded into the unmodified reference (in the dev container, by
``oracle/make_golden.py``) and into this repo's modules anywhere else.  Unlike the
reference's own init, biases are non-zero so that bias handling is actually exercised.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Sequence, Tuple

import torch

Tensor = torch.Tensor


def decoder_param_shapes(n_views: int = 3, cos_n_group: Sequence[int] = (2, 8), W: int = 128, D: int = 6,
                         skips: Sequence[int] = (4,), L_3D: int = 10) -> "OrderedDict[str, Tuple[int, ...]]":
    """Parameter tree of the reference ``nerf_dec`` (models/rfdecoder/cond_nerf.py:15-50)."""
    in3d = 3 + 6 * L_3D
    in_feat = sum(cos_n_group) + n_views * 4
    sh: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    sh["pts_linears.0.weight"] = (W, in3d)
    sh["pts_linears.0.bias"] = (W,)
    for i in range(D - 1):
        k = W + in3d if i in skips else W
        sh[f"pts_linears.{i + 1}.weight"] = (W, k)
        sh[f"pts_linears.{i + 1}.bias"] = (W,)
    sh["pts_bias.weight"] = (W, in_feat)
    sh["pts_bias.bias"] = (W,)
    sh["views_linears.0.weight"] = (W // 2, W + 3)
    sh["views_linears.0.bias"] = (W // 2,)
    sh["alpha_linear.0.weight"] = (16, W)
    sh["alpha_linear.0.bias"] = (16,)
    for n in ("w_qs", "w_ks", "w_vs", "fc"):
        sh[f"ray_attention.{n}.weight"] = (16, 16)
    sh["ray_attention.layer_norm.weight"] = (16,)
    sh["ray_attention.layer_norm.bias"] = (16,)
    sh["out_alpha_linear.0.weight"] = (16, 16)
    sh["out_alpha_linear.0.bias"] = (16,)
    sh["out_alpha_linear.2.weight"] = (1, 16)
    sh["out_alpha_linear.2.bias"] = (1,)
    sh["feature_linear.weight"] = (W, W)
    sh["feature_linear.bias"] = (W,)
    sh["rgb_linear.weight"] = (3, W // 2)
    sh["rgb_linear.bias"] = (3,)
    return sh


def encoder_param_shapes(n_layers: int = 6, C: int = 128, n_up_blocks: int = 1) -> "OrderedDict[str, Tuple[int, ...]]":
    """Parameter tree of the reference ``feat_enc`` (models/gmflow/{backbone,transformer,superres}.py)."""
    sh: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    sh["backbone.conv1.weight"] = (64, 3, 7, 7)
    cin = 64
    for name, dim, stride in (("layer1", 64, 1), ("layer2", 96, 2), ("layer3", 128, 2)):
        for b in range(2):
            ci = cin if b == 0 else dim
            st = stride if b == 0 else 1
            sh[f"backbone.{name}.{b}.conv1.weight"] = (dim, ci, 3, 3)
            sh[f"backbone.{name}.{b}.conv2.weight"] = (dim, dim, 3, 3)
            if st != 1 or ci != dim:
                sh[f"backbone.{name}.{b}.downsample.0.weight"] = (dim, ci, 1, 1)
                sh[f"backbone.{name}.{b}.downsample.0.bias"] = (dim,)
        cin = dim
    sh["backbone.conv2.weight"] = (C, 128, 1, 1)
    sh["backbone.conv2.bias"] = (C,)
    for i in range(n_layers):
        for blk, ffn in (("self_attn", False), ("cross_attn_ffn", True)):
            p = f"transformer.layers.{i}.{blk}."
            for n in ("q_proj", "k_proj", "v_proj", "merge"):
                sh[p + n + ".weight"] = (C, C)
            sh[p + "norm1.weight"] = (C,)
            sh[p + "norm1.bias"] = (C,)
            if ffn:
                sh[p + "mlp.0.weight"] = (8 * C, 2 * C)
                sh[p + "mlp.2.weight"] = (C, 8 * C)
                sh[p + "norm2.weight"] = (C,)
                sh[p + "norm2.bias"] = (C,)
    for i in range(n_up_blocks):
        sh[f"featup_net.conv_ls.{i}.weight"] = (C, C, 3, 3)
        sh[f"featup_net.conv_ls.{i}.bias"] = (C,)
    for i in range(n_up_blocks + 1):
        sh[f"featup_net.conv_l2rs.{i}.weight"] = (C, C, 3, 3)
        sh[f"featup_net.conv_l2rs.{i}.bias"] = (C,)
    return sh


def synthetic_state_dict(shapes: "OrderedDict[str, Tuple[int, ...]]", seed: int = 0,
                         weight_gain: float = 1.4, bias_std: float = 0.05) -> "OrderedDict[str, Tensor]":
    """One N(0, gain^2/fan_in) tensor per weight, small non-zero biases, norm scales near 1."""
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        if "norm" in name and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = bias_std * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) * (weight_gain / math.sqrt(fan_in))
        sd[name] = t.to(torch.float32)
    return sd


def synthetic_decoder(seed: int = 0, density_gain: float = 1.0, **kw) -> "OrderedDict[str, Tensor]":
    """Decoder weights whose density head is neither dead nor saturated (SURVEY 8c degeneracy warning)."""
    sd = synthetic_state_dict(decoder_param_shapes(**kw), seed)
    # keep sigma = relu(w . a1 + b) alive and O(0.01) (opacity ~0.5 at S=64, ~40% of samples at the ReLU floor)
    sd["out_alpha_linear.2.bias"] = torch.tensor([0.01 * density_gain])
    sd["out_alpha_linear.2.weight"] = sd["out_alpha_linear.2.weight"] * (0.05 * density_gain)
    # the conditioning gate multiplies every trunk layer; keep it O(1) so the trunk neither dies nor explodes
    sd["pts_bias.bias"] = sd["pts_bias.bias"] + 0.9
    return sd


def synthetic_encoder(seed: int = 0, **kw) -> "OrderedDict[str, Tensor]":
    return synthetic_state_dict(encoder_param_shapes(**kw), seed, weight_gain=1.0)


# ----------------------------------------------------------------------------
# synthetic scene (cameras follow SURVEY.md 8c: DTU-like intrinsics scaled to the image size)
# ----------------------------------------------------------------------------
def _w2c(deg_y: float, deg_x: float = 0.0, dist: float = 3.3) -> Tensor:
    a, b = math.radians(deg_y), math.radians(deg_x)
    Ry = torch.tensor([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
    Rx = torch.tensor([[1, 0, 0], [0, math.cos(b), -math.sin(b)], [0, math.sin(b), math.cos(b)]])
    E = torch.eye(4)
    E[:3, :3] = Rx @ Ry
    E[2, 3] = dist
    return E


def synthetic_cameras(H: int, W: int, baseline_deg: float = 10.0, near: float = 2.125, far: float = 4.525):
    """3 source views at (-b, 0, +b) degrees and a target at (3, 2) degrees; intrinsics scaled from
    the 512x640 DTU-like K of SURVEY.md 8c.  Returns extrinsics [1,4,4,4], intrinsics [1,4,3,3],
    near_fars [1,4,2] in the reference batch schema (last entry = target)."""
    extr = torch.stack([_w2c(-baseline_deg), _w2c(0.0), _w2c(baseline_deg), _w2c(3.0, 2.0)])[None]
    sx, sy = W / 640.0, H / 512.0
    K = torch.tensor([[1446.2 * sx, 0, 320.0 * sx], [0, 1441.6 * sy, 256.0 * sy], [0, 0, 1]])
    intr = K[None, None].repeat(1, 4, 1, 1)
    nf = torch.tensor([near, far])[None, None].repeat(1, 4, 1)
    return extr, intr, nf


def synthetic_scene(H: int, W: int, seed: int = 1234, V: int = 3, C: int = 256):
    """Random feature maps (reference layout [1,V,C,h,w]) at 1/8 and 1/4 scale + images [1,V,3,H,W]."""
    g = torch.Generator().manual_seed(seed)
    f8 = torch.randn(1, V, C, H // 8, W // 8, generator=g)
    f4 = torch.randn(1, V, C, H // 4, W // 4, generator=g)
    imgs = torch.rand(1, V, 3, H, W, generator=g)
    return [f8, f4], imgs, g
