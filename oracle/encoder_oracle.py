"""fp32 CPU restatement of the GMFlow feature encoder (oracle; test infrastructure only).

The hot-path piece is the split-window single-head attention
(models/gmflow/transformer.py:8-105); the surrounding CNN / linear / FFN /
upsampler layers are restated functionally on a ``feat_enc`` state_dict so the
whole encoder can be checked end to end.  Window attention is written here with
explicit index arithmetic (no ``torch.roll``, no materialised mask tensor) so it
is an independent statement of what the reference computes.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# window attention (the K-attn oracle)
# ----------------------------------------------------------------------------
def shift_region_id(pos: Tensor, size: int, win: int, shift: int) -> Tensor:
    """Region label (0,1,2) of *rolled* coordinate ``pos`` along one axis.

    models/gmflow/transformer.py:25-36: slices [0,-win), [-win,-shift), [-shift,end).
    """
    return (pos >= size - win).long() + (pos >= size - shift).long()


def window_attention(q: Tensor, k: Tensor, v: Tensor, h: int, w: int, num_splits: int, with_shift: bool) -> Tensor:
    """q,k,v [B, h*w, C] -> [B, h*w, C].

    ``num_splits == 1``: full attention, models/gmflow/transformer.py:8-16.
    Otherwise models/gmflow/transformer.py:46-105: roll by (-wh/2, -ww/2) when
    shifted, partition into num_splits^2 windows, ``softmax(q k^T / sqrt(C) +
    mask)`` with mask -100 between different shift regions (:19-43), ``. v``,
    merge, roll back.
    """
    B, L, C = q.shape
    scale = 1.0 / math.sqrt(C)
    if num_splits == 1:
        p = torch.softmax((q @ k.transpose(1, 2)) * scale, dim=-1)
        return p @ v
    wh, ww = h // num_splits, w // num_splits
    sh, sw = (wh // 2, ww // 2) if with_shift else (0, 0)
    out = torch.empty_like(q)
    yy, xx = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
    for wy in range(num_splits):
        for wx in range(num_splits):
            # coordinates of this window's tokens in the rolled frame, then in the original frame
            ry = (wy * wh + yy).reshape(-1)
            rx = (wx * ww + xx).reshape(-1)
            oy = (ry + sh) % h
            ox = (rx + sw) % w
            tok = oy * w + ox
            qs, ks, vs = q[:, tok], k[:, tok], v[:, tok]
            s = (qs @ ks.transpose(1, 2)) * scale
            if with_shift:
                reg = shift_region_id(ry, h, wh, sh) * 3 + shift_region_id(rx, w, ww, sw)
                s = s + torch.where(reg[:, None] != reg[None, :], -100.0, 0.0)[None]
            out[:, tok] = torch.softmax(s, dim=-1) @ vs
    return out


# ----------------------------------------------------------------------------
# the rest of the encoder, functional on a state_dict
# ----------------------------------------------------------------------------
def _ln(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def post_attention(sd: Dict[str, Tensor], pfx: str, src: Tensor, msg: Tensor, ffn: bool, quant=None) -> Tensor:
    """models/gmflow/transformer.py:173-185: merge, norm1, (mlp on cat[source, message], norm2), residual.  ``quant`` (tests of
    the fp16-operand kernel): a function applied to every GEMM operand (activations and weights) before the product."""
    qz = quant if quant is not None else (lambda t: t)
    msg = _ln(qz(msg) @ qz(sd[pfx + "merge.weight"]).T, sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"])
    if ffn:
        x = qz(torch.cat([src, msg], dim=-1)) @ qz(sd[pfx + "mlp.0.weight"]).T
        x = 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))          # exact GELU
        msg = _ln(qz(x) @ qz(sd[pfx + "mlp.2.weight"]).T, sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"])
    return src + msg


def transformer_layer(sd: Dict[str, Tensor], pfx: str, src: Tensor, tgt: Tensor, h: int, w: int,
                      num_splits: int, with_shift: bool, ffn: bool) -> Tensor:
    """models/gmflow/transformer.py:147-185."""
    q = src @ sd[pfx + "q_proj.weight"].T
    k = tgt @ sd[pfx + "k_proj.weight"].T
    v = tgt @ sd[pfx + "v_proj.weight"].T
    msg = window_attention(q, k, v, h, w, num_splits, with_shift)
    return post_attention(sd, pfx, src, msg, ffn)


def feature_transformer(sd: Dict[str, Tensor], f0: Tensor, f1: Tensor, num_splits: int, n_layers: int,
                        wo_self_attn: bool = False) -> Tuple[Tensor, Tensor]:
    """f0,f1 [P, C, h, w] -> same.  models/gmflow/transformer.py:279-339 (+ block :216-247).

    Both directions are batched ([f0;f1] attends to [f1;f0]); the cross-attention
    target of block i is the other view's output of block i-1; odd blocks are shifted.
    """
    P, C, h, w = f0.shape
    a = f0.flatten(2).permute(0, 2, 1)
    b = f1.flatten(2).permute(0, 2, 1)
    x = torch.cat([a, b], 0)
    y = torch.cat([b, a], 0)
    for i in range(n_layers):
        shift = (i % 2 == 1) and num_splits > 1
        pfx = f"transformer.layers.{i}."
        if not wo_self_attn:
            x = transformer_layer(sd, pfx + "self_attn.", x, x, h, w, num_splits, shift, ffn=False)
        x = transformer_layer(sd, pfx + "cross_attn_ffn.", x, y, h, w, num_splits, shift, ffn=True)
        y = torch.cat([x[P:], x[:P]], 0)
    o0 = x[:P].reshape(P, h, w, C).permute(0, 3, 1, 2).contiguous()
    o1 = x[P:].reshape(P, h, w, C).permute(0, 3, 1, 2).contiguous()
    return o0, o1


def sine_position(hw: int, ww: int, C: int) -> Tensor:
    """[C, hw, ww] DETR sine embedding.  models/gmflow/position.py:26-47 (normalised, scale 2*pi)."""
    npf = C // 2
    yv = torch.arange(1, hw + 1, dtype=torch.float32) / (hw + 1e-6) * (2 * math.pi)
    xv = torch.arange(1, ww + 1, dtype=torch.float32) / (ww + 1e-6) * (2 * math.pi)
    i = torch.arange(npf, dtype=torch.float32)
    dim_t = 10000.0 ** (2 * torch.div(i, 2, rounding_mode="trunc") / npf)
    px = xv[:, None] / dim_t
    py = yv[:, None] / dim_t
    px = torch.stack([px[:, 0::2].sin(), px[:, 1::2].cos()], dim=2).flatten(1)     # [ww, npf]
    py = torch.stack([py[:, 0::2].sin(), py[:, 1::2].cos()], dim=2).flatten(1)     # [hw, npf]
    pos = torch.cat([py[:, None, :].expand(hw, ww, npf), px[None, :, :].expand(hw, ww, npf)], dim=-1)
    return pos.permute(2, 0, 1).contiguous()


def add_window_position(f: Tensor, num_splits: int) -> Tensor:
    """Add the sine embedding computed per window.  models/gmflow/utils.py:68-88."""
    P, C, h, w = f.shape
    pos = sine_position(h // num_splits, w // num_splits, C)
    return f + pos.repeat(1, num_splits, num_splits)[None]


def _inorm(x: Tensor) -> Tensor:
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5)


def _res_block(sd: Dict[str, Tensor], pfx: str, x: Tensor, stride: int) -> Tensor:
    """models/gmflow/backbone.py:6-36."""
    y = torch.relu(_inorm(F.conv2d(x, sd[pfx + "conv1.weight"], None, stride, 1)))
    y = torch.relu(_inorm(F.conv2d(y, sd[pfx + "conv2.weight"], None, 1, 1)))
    if (pfx + "downsample.0.weight") in sd:
        x = _inorm(F.conv2d(x, sd[pfx + "downsample.0.weight"], sd[pfx + "downsample.0.bias"], stride, 0))
    return torch.relu(x + y)


def cnn_backbone(sd: Dict[str, Tensor], img: Tensor) -> Tensor:
    """[n,3,H,W] (ImageNet-normalised) -> [n,128,H/8,W/8].  models/gmflow/backbone.py:101-122."""
    x = torch.relu(_inorm(F.conv2d(img, sd["backbone.conv1.weight"], None, 2, 3)))
    for name, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
        x = _res_block(sd, f"backbone.{name}.0.", x, stride)
        x = _res_block(sd, f"backbone.{name}.1.", x, 1)
    return F.conv2d(x, sd["backbone.conv2.weight"], sd["backbone.conv2.bias"])


def upsampler(sd: Dict[str, Tensor], x: Tensor, n_blocks: int) -> Tensor:
    """models/gmflow/superres.py:26-38."""
    right = F.conv2d(x, sd["featup_net.conv_l2rs.0.weight"], sd["featup_net.conv_l2rs.0.bias"], 1, 1)
    left = x
    for i in range(n_blocks):
        left = F.interpolate(left, scale_factor=2.0, mode="nearest")
        left = F.leaky_relu(F.conv2d(left, sd[f"featup_net.conv_ls.{i}.weight"], sd[f"featup_net.conv_ls.{i}.bias"], 1, 1), 0.2)
        mid = F.conv2d(left, sd[f"featup_net.conv_l2rs.{i + 1}.weight"], sd[f"featup_net.conv_l2rs.{i + 1}.bias"], 1, 1)
        right = F.interpolate(right, scale_factor=2.0, mode="bilinear", align_corners=False) + mid
    return right


def encode_views(sd: Dict[str, Tensor], images: Tensor, num_splits: int = 2, n_layers: int = 6,
                 upsample_factor: int = 2, wo_self_attn: bool = False) -> List[Tensor]:
    """images [V,3,H,W] in [0,1] -> [feat_1/8 [V,256,h,w], feat_1/4 [V,256,2h,2w]] (reference NCHW).

    models/gmflow/gmflow.py:47-150 + models/matchnerf.py:183-207: backbone on
    every view, the V(V-1)/2 ordered pairs through the transformer, optional
    upsampler, then regroup so view i holds [f(i | partner_a), f(i | partner_b)].
    """
    V = images.shape[0]
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    base = cnn_backbone(sd, (images - mean) / std)
    pairs = [(a, b) for a in range(V - 1) for b in range(a + 1, V)]
    f0 = torch.stack([base[a] for a, _ in pairs])
    f1 = torch.stack([base[b] for _, b in pairs])
    f0 = add_window_position(f0, num_splits)
    f1 = add_window_position(f1, num_splits)
    f0, f1 = feature_transformer(sd, f0, f1, num_splits, n_layers, wo_self_attn)
    n_blocks = int(math.log2(upsample_factor))
    up = upsampler(sd, torch.cat([f0, f1], 0), n_blocks)
    u0, u1 = up[: len(pairs)], up[len(pairs):]
    outs = []
    for a0, a1 in ((f0, f1), (u0, u1)):
        per_view = [[] for _ in range(V)]
        for p, (i, j) in enumerate(pairs):
            per_view[i].append(a0[p])
            per_view[j].append(a1[p])
        outs.append(torch.stack([torch.cat(x, 0) for x in per_view]))
    return outs
