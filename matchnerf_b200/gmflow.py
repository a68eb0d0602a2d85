"""Host-side mirror of the reference feature encoder (``models/gmflow``): same module tree and
parameter names, so ``feat_enc`` checkpoints (and GMFlow pre-training weights) load unchanged.

What runs where
  * split-window attention (models/gmflow/transformer.py:46-105, :8-16): this repo's CUDA kernel through the
    C ABI (``mnf_window_attn_fwd``) -- no roll, no window copies, no L x L score tensor;
  * the CNN backbone, linear projections, LayerNorm, FFN and the up-sampler: PyTorch library calls
    (cuDNN / cuBLAS), as SURVEY.md 8(f) schedules them after the per-ray path.
There is no CPU execution path for the attention: off-GPU the forward raises.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import capi


# ----------------------------------------------------------------------------------------------
# CNN backbone (models/gmflow/backbone.py) -- parameter names: conv1, layer{1,2,3}.{0,1}.{conv1,conv2,
# downsample.0}, conv2
# ----------------------------------------------------------------------------------------------
def _in_relu(x, mode: int, residual=None):
    """InstanceNorm2d (no affine) fused with what follows it in the backbone: mode 0 = norm only, 1 = norm + ReLU,
    2 = relu(residual + relu(norm(x))).  On a CUDA tensor (inference) this is one kernel of this repo's library
    (mnf_instance_norm_fwd); elsewhere (CPU tensors, autograd) the PyTorch ops the reference uses."""
    if x.is_cuda and not (torch.is_grad_enabled() and x.requires_grad) and x.dtype == torch.float32:
        return capi.get_context(x.device).instance_norm(x.contiguous(), mode, residual.contiguous() if residual is not None else None)
    y = F.instance_norm(x)
    if mode >= 1:
        y = F.relu(y)
    if mode == 2:
        y = F.relu(residual + y)
    return y


class ResidualBlock(nn.Module):
    """Two 3x3 convs with instance norm + identity/1x1 shortcut (models/gmflow/backbone.py:6-36)."""

    def __init__(self, c_in: int, c_out: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(c_in, c_out, 3, stride, 1, bias=False)
        self.conv2 = nn.Conv2d(c_out, c_out, 3, 1, 1, bias=False)
        self.downsample = None
        if stride != 1 or c_in != c_out:
            self.downsample = nn.Sequential(nn.Conv2d(c_in, c_out, 1, stride), nn.InstanceNorm2d(c_out))

    def forward(self, x):
        y = _in_relu(self.conv1(x), 1)
        y2 = self.conv2(y)
        if self.downsample is not None:
            x = _in_relu(self.downsample[0](x), 0)
        return _in_relu(y2, 2, x)                 # relu(x + relu(IN(conv2(y))))


class CNNEncoder(nn.Module):
    """1/8-resolution 128-channel features (models/gmflow/backbone.py:39-122, ``num_output_scales = 1``)."""

    def __init__(self, output_dim: int = 128):
        super().__init__()
        dims = (64, 96, 128)
        self.conv1 = nn.Conv2d(3, dims[0], 7, 2, 3, bias=False)
        self.layer1 = nn.Sequential(ResidualBlock(dims[0], dims[0], 1), ResidualBlock(dims[0], dims[0], 1))
        self.layer2 = nn.Sequential(ResidualBlock(dims[0], dims[1], 2), ResidualBlock(dims[1], dims[1], 1))
        self.layer3 = nn.Sequential(ResidualBlock(dims[1], dims[2], 2), ResidualBlock(dims[2], dims[2], 1))
        self.conv2 = nn.Conv2d(dims[2], output_dim, 1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    # Inference on the GPU runs the convolutions as cuDNN channels-last kernels (no NCHW <-> NHWC conversion kernels around them)
    # and the instance norms as this repo's NHWC kernel (mnf_instance_norm_nhwc_fwd).  torch.float32 (default): TF32 convolutions,
    # fp32 activations -- the same arithmetic as the NCHW path.  torch.float16: fp16 activations as well (0.2 ms faster at DTU
    # size, feature maps 2.2e-3 instead of 1.7e-3 relative RMS from an all-fp32 run: measured, tools/r02_feat_err.py; not the
    # default).  None: fp32 NCHW convolutions + the per-plane fp32 kernel.
    fast_dtype = torch.float32

    def forward(self, x):
        if self.fast_dtype is not None and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled():
            return self._forward_nhwc(x)
        x = _in_relu(self.conv1(x), 1)
        x = self.layer3(self.layer2(self.layer1(x)))
        return self.conv2(x)

    def _nhwc_params(self):
        key = (self.fast_dtype,) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if getattr(self, "_f16_cache", None) is None or self._f16_cache[0] != key:
            conv = {}
            for name, m in self.named_modules():
                if isinstance(m, nn.Conv2d):
                    conv[name] = (m.weight.detach().to(self.fast_dtype).contiguous(memory_format=torch.channels_last),
                                  None if m.bias is None else m.bias.detach().to(self.fast_dtype), m.stride, m.padding)
            self._f16_cache = (key, conv)
        return self._f16_cache[1]

    def _forward_nhwc(self, x, keep_nhwc: bool = False):
        ctx = capi.get_context(x.device)
        P = self._nhwc_params()

        def conv(h, name):
            w, b, stride, pad = P[name]
            return F.conv2d(h, w, b, stride, pad)
        h = x.to(self.fast_dtype).contiguous(memory_format=torch.channels_last)
        h = ctx.instance_norm_nhwc(conv(h, "conv1"), 1)
        for li in (1, 2, 3):
            for bi in (0, 1):
                pre = f"layer{li}.{bi}"
                y = ctx.instance_norm_nhwc(conv(h, pre + ".conv1"), 1)
                y2 = conv(y, pre + ".conv2")
                if (pre + ".downsample.0") in P:
                    h = ctx.instance_norm_nhwc(conv(h, pre + ".downsample.0"), 0)
                h = ctx.instance_norm_nhwc(y2, 2, h)             # relu(x + relu(IN(conv2(y))))
        out = conv(h, "conv2").float()
        return out if keep_nhwc else out.contiguous()      # keep_nhwc: channels_last strides (the token layout of the transformer)


# ----------------------------------------------------------------------------------------------
# transformer (models/gmflow/transformer.py) -- names: layers.{i}.{self_attn,cross_attn_ffn}.{q_proj,k_proj,
# v_proj,merge,norm1[,mlp.0,mlp.2,norm2]}
# ----------------------------------------------------------------------------------------------
def window_attention(q, k, v, h: int, w: int, num_splits: int, with_shift: bool, impl: int = 0):
    """Dispatch to the CUDA kernel behind the C ABI.  q, k, v: [B, h*w, 128] fp32 on a CUDA device."""
    if not q.is_cuda:
        raise RuntimeError("matchnerf_b200: split-window attention only exists as a CUDA kernel (no CPU path)")
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
        from .train_path import window_attention_autograd        # training step: index-gathered batched GEMMs + autograd
        return window_attention_autograd(q, k, v, h, w, num_splits, with_shift)
    return capi.get_context(q.device).window_attn(q, k, v, h, w, num_splits, with_shift, impl)


class TransformerLayer(nn.Module):
    """models/gmflow/transformer.py:108-185."""

    def __init__(self, d_model: int = 128, no_ffn: bool = False, ffn_dim_expansion: int = 4, with_shift: bool = False):
        super().__init__()
        self.no_ffn, self.with_shift = no_ffn, with_shift
        self.q_proj = nn.Linear(d_model, d_model, bias=False)
        self.k_proj = nn.Linear(d_model, d_model, bias=False)
        self.v_proj = nn.Linear(d_model, d_model, bias=False)
        self.merge = nn.Linear(d_model, d_model, bias=False)
        self.norm1 = nn.LayerNorm(d_model)
        if not no_ffn:
            c = 2 * d_model
            self.mlp = nn.Sequential(nn.Linear(c, c * ffn_dim_expansion, bias=False), nn.GELU(),
                                     nn.Linear(c * ffn_dim_expansion, d_model, bias=False))
            self.norm2 = nn.LayerNorm(d_model)

    def forward(self, source, target, height: int, width: int, attn_num_splits: int, target_roll: int = 0):
        """``target_roll`` r (not in the reference): keys / values of batch item b come from ``target[(b + r) % B]`` -- lets
        FeatureTransformer pair every view with the other one without materialising ``cat([x[b:], x[:b]])``."""
        shift = self.with_shift and attn_num_splits > 1
        if target_roll and not (self.fused_proj and self._fused_norms(source)):
            target, target_roll = torch.roll(target, -target_roll, 0), 0
        if self.fused_proj and self._fused_norms(source):
            # inference on the GPU: the three projections run inside the kernel that builds the attention operands
            # (mnf_window_attn_proj_fwd): no fp32 q / k / v tensors, no pre-pack pass
            ctx = capi.get_context(source.device)
            msg = ctx.window_attn_proj(source, target, self._proj_weights(ctx), height, width, attn_num_splits, shift, target_roll)
        else:
            msg = window_attention(self.q_proj(source), self.k_proj(target), self.v_proj(target), height, width,
                                   attn_num_splits, shift)
        if self.fused_block and self._fused_norms(msg):
            # inference on the GPU: merge + LayerNorm (+ FFN + LayerNorm) + residual = ONE tcgen05 kernel of this repo's library
            # (mnf_token_block_fwd); the 1024-wide hidden tensor never leaves the SM
            ctx = capi.get_context(msg.device)
            return ctx.token_block(msg, source, self._block_weights(ctx), not self.no_ffn, self.norm1.eps)
        msg = self.merge(msg)
        if not self._fused_norms(msg):                       # CPU tensors / autograd: the PyTorch ops the reference uses
            msg = self.norm1(msg)
            if not self.no_ffn:
                msg = self.norm2(self._ffn(torch.cat([source, msg], dim=-1)))
            return source + msg
        # inference on the GPU: every LayerNorm is one kernel of this repo's library together with the add / cat / dtype conversion
        # around it (mnf_token_layernorm_fwd)
        ctx = capi.get_context(msg.device)
        if self.no_ffn:
            return ctx.token_layernorm(msg, self.norm1.weight, self.norm1.bias, self.norm1.eps, residual=source)       # source + LN1(.)
        dt = self.ffn_dtype
        w1, w2 = self._ffn_weights(dt)
        x16 = ctx.token_layernorm(msg, self.norm1.weight, self.norm1.bias, self.norm1.eps, prefix=source)             # half(cat[source, LN1(.)])
        hid = F.linear(F.gelu(F.linear(x16, w1)), w2)                                                                # fp16, fp32 accumulation
        return ctx.token_layernorm(hid, self.norm2.weight, self.norm2.bias, self.norm2.eps, residual=source)         # source + LN2(.)

    # dtype of the FFN's operands and 1024-wide hidden tensor.  The TF32 path (None) already rounds every GEMM operand to a
    # 10-bit mantissa; storing the hidden activations as fp16 (same mantissa, fp32 accumulation in both GEMMs, GELU
    # evaluated in fp32 inside the kernel) halves the 126 MB-per-layer hidden round trips.  Inference only.
    ffn_dtype = torch.float16
    # merge / LayerNorm / FFN / residual as one kernel (csrc/token_block.cu); False = the library-GEMM + token_layernorm path (A/B runs)
    fused_block = os.environ.get("MNF_FUSED_BLOCK", "1") != "0"

    fused_proj = os.environ.get("MNF_FUSED_PROJ", "1") != "0"     # q / k / v projections inside the attention operand-packing kernel

    def _proj_weights(self, ctx):
        ps = [self.q_proj.weight, self.k_proj.weight, self.v_proj.weight]
        key = (ctx.device,) + tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, "_proj_cache", None) is None or self._proj_cache[0] != key:
            self._proj_cache = (key, ctx.window_attn_pack_proj(*ps))
        return self._proj_cache[1]

    def _block_weights(self, ctx):
        ps = [self.merge.weight, self.norm1.weight, self.norm1.bias]
        if not self.no_ffn:
            ps += [self.mlp[0].weight, self.mlp[2].weight, self.norm2.weight, self.norm2.bias]
        key = (ctx.device,) + tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, "_block_cache", None) is None or self._block_cache[0] != key:
            if not self.no_ffn and self.norm2.eps != self.norm1.eps:
                raise NotImplementedError("token_block assumes one LayerNorm eps per layer")
            self._block_cache = (key, ctx.token_block_pack(*ps))
        return self._block_cache[1]

    def _fused_norms(self, x) -> bool:
        return (x.is_cuda and x.dtype == torch.float32 and x.shape[-1] == 128 and self.ffn_dtype == torch.float16
                and not (torch.is_grad_enabled() and x.requires_grad))

    def _ffn(self, x):
        dt = self.ffn_dtype
        if dt is None or not x.is_cuda or (torch.is_grad_enabled() and x.requires_grad):
            return self.mlp(x)
        w1, w2 = self._ffn_weights(dt)
        return F.linear(F.gelu(F.linear(x.to(dt), w1)), w2).float()

    def _ffn_weights(self, dt):
        key = (dt, self.mlp[0].weight._version, self.mlp[2].weight._version, self.mlp[0].weight.data_ptr())
        if getattr(self, "_ffn_cache", None) is None or self._ffn_cache[0] != key:
            self._ffn_cache = (key, self.mlp[0].weight.detach().to(dt), self.mlp[2].weight.detach().to(dt))
        return self._ffn_cache[1], self._ffn_cache[2]


class TransformerBlock(nn.Module):
    """Self-attention (no FFN) then cross-attention + FFN (models/gmflow/transformer.py:188-247)."""

    def __init__(self, d_model: int = 128, ffn_dim_expansion: int = 4, with_shift: bool = False):
        super().__init__()
        self.self_attn = TransformerLayer(d_model, True, ffn_dim_expansion, with_shift)
        self.cross_attn_ffn = TransformerLayer(d_model, False, ffn_dim_expansion, with_shift)

    def forward(self, source, target, height, width, attn_num_splits, wo_self_attn=False, target_roll: int = 0):
        if not wo_self_attn:
            source = self.self_attn(source, source, height, width, attn_num_splits)
        return self.cross_attn_ffn(source, target, height, width, attn_num_splits, target_roll)


class FeatureTransformer(nn.Module):
    """models/gmflow/transformer.py:250-339."""

    def __init__(self, num_layers: int = 6, d_model: int = 128, ffn_dim_expansion: int = 4):
        super().__init__()
        self.d_model = d_model
        self.layers = nn.ModuleList([TransformerBlock(d_model, ffn_dim_expansion, with_shift=(i % 2 == 1))
                                     for i in range(num_layers)])
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, feature0, feature1, attn_num_splits: int, wo_self_attn: bool = False):
        b, c, h, w = feature0.shape
        x = torch.cat([feature0, feature1], 0).flatten(2).transpose(1, 2).contiguous()      # [2P, hw, C]
        for layer in self.layers:
            y = torch.cat([x[b:], x[:b]], 0)          # the other view's tokens from before this block (:331)
            x = layer(x, y, h, w, attn_num_splits, wo_self_attn)
        x = x.transpose(1, 2).reshape(2 * b, c, h, w)
        return x[:b].contiguous(), x[b:].contiguous()

    def forward_tokens(self, x, b: int, h: int, w: int, attn_num_splits: int, wo_self_attn: bool = False):
        """The same layers on tokens: x [2b, h*w, C] (first b items = feature0 of the pairs, last b = feature1) -> [2b, h*w, C].
        The other view's tokens from before the block (:331) are addressed by a batch roll inside the attention kernel."""
        for layer in self.layers:
            x = layer(x, x, h, w, attn_num_splits, wo_self_attn, target_roll=b)
        return x


def sine_position(hw: int, ww: int, channels: int, device, dtype=torch.float32):
    """[C, hw, ww] DETR sine embedding, normalised to 2*pi (models/gmflow/position.py:26-47)."""
    npf = channels // 2
    ys = torch.arange(1, hw + 1, dtype=torch.float32, device=device) / (hw + 1e-6) * (2 * math.pi)
    xs = torch.arange(1, ww + 1, dtype=torch.float32, device=device) / (ww + 1e-6) * (2 * math.pi)
    i = torch.arange(npf, dtype=torch.float32, device=device)
    dim_t = 10000.0 ** (2 * torch.div(i, 2, rounding_mode="trunc") / npf)

    def interleave(v):
        a = v[:, None] / dim_t
        return torch.stack([a[:, 0::2].sin(), a[:, 1::2].cos()], dim=2).flatten(1)

    py, px = interleave(ys), interleave(xs)
    pos = torch.cat([py[:, None, :].expand(hw, ww, npf), px[None, :, :].expand(hw, ww, npf)], dim=-1)
    return pos.permute(2, 0, 1).contiguous().to(dtype)


# ----------------------------------------------------------------------------------------------
# up-sampler (models/gmflow/superres.py) -- names: conv_ls.{i}, conv_l2rs.{i}
# ----------------------------------------------------------------------------------------------
class UpSampler(nn.Module):
    def __init__(self, n_feat: int = 128, upsample_factor: int = 2):
        super().__init__()
        self.n_blocks = int(math.log2(upsample_factor))
        self.conv_ls = nn.ModuleList([nn.Conv2d(n_feat, n_feat, 3, 1, 1) for _ in range(self.n_blocks)])
        self.conv_l2rs = nn.ModuleList([nn.Conv2d(n_feat, n_feat, 3, 1, 1) for _ in range(self.n_blocks + 1)])

    def forward(self, x):
        right, left = self.conv_l2rs[0](x), x
        for i in range(self.n_blocks):
            left = F.leaky_relu(self.conv_ls[i](F.interpolate(left, scale_factor=2.0, mode="nearest")), 0.2)
            right = F.interpolate(right, scale_factor=2.0, mode="bilinear", align_corners=False) + self.conv_l2rs[i + 1](left)
        return right


# ----------------------------------------------------------------------------------------------
# GMFlow wrapper (models/gmflow/gmflow.py)
# ----------------------------------------------------------------------------------------------
class _matmul_precision:
    """Scoped cuBLAS/cuDNN math mode for the encoder's library calls.  "tf32" lets the 128/256/1024-wide linear layers
    (48 + 145 GFLOP of fp32 SGEMM per DTU triplet) run on the tensor cores with 10-bit-mantissa operands and
    fp32 accumulation -- the same operand precision the decoder kernel uses; "fp32" keeps PyTorch's default."""

    def __init__(self, mode: str):
        self.mode = mode

    def __enter__(self):
        self.prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        if self.mode == "tf32":
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.prev
        return False


class GMFlow(nn.Module):
    """Encoder used by MatchNeRF: backbone -> pairwise transformer -> up-sampler.

    ``forward`` keeps the reference's return convention (``aug_feat0s`` / ``aug_feat1s`` lists holding the raw
    1/8 features then the up-sampled ones, each [B, P, 128, h, w]); models/gmflow/gmflow.py:91-150.
    """

    def __init__(self, num_scales=1, upsample_factor=2, feature_channels=128, attention_type="swin",
                 num_transformer_layers=6, ffn_dim_expansion=4, num_head=1, feature_upsampler="network",
                 device="cuda", **kwargs):
        super().__init__()
        if num_scales != 1 or attention_type != "swin" or num_head != 1:
            raise NotImplementedError("only the configuration MatchNeRF instantiates (1 scale, swin, 1 head) is built")
        self.feature_channels = feature_channels
        self.feature_upsampler = feature_upsampler
        self.backbone = CNNEncoder(feature_channels)
        self.transformer = FeatureTransformer(num_transformer_layers, feature_channels, ffn_dim_expansion)
        if feature_upsampler == "network":
            self.featup_net = UpSampler(feature_channels, upsample_factor)

    _norm_consts = {}

    @staticmethod
    def normalize_images(images):
        """ImageNet mean / std normalisation (models/gmflow/gmflow.py:82-89).  The two constant tensors are created once per
        device: a fresh torch.tensor(...) per call is a pageable host->device copy in the middle of the launch stream."""
        c = GMFlow._norm_consts.get(images.device)
        if c is None:
            c = (torch.tensor([0.485, 0.456, 0.406], device=images.device).view(1, 1, 3, 1, 1),
                 torch.tensor([0.229, 0.224, 0.225], device=images.device).view(1, 1, 3, 1, 1))
            GMFlow._norm_consts[images.device] = c
        return (images - c[0]) / c[1]

    matmul_precision = "tf32"      # "tf32" | "fp32" for the cuBLAS / cuDNN calls of the encoder
    # memory formats of the cuDNN convolution stacks (measured on B200, DTU size, tools/prof_encoder.py): the up-sampler
    # is 0.87 -> 0.51 ms in channels_last (no NCHW<->NHWC copies around every conv); the backbone is not faster (its
    # instance norms dominate), so it stays contiguous
    backbone_memory_format = torch.contiguous_format
    upsampler_memory_format = torch.channels_last

    def forward(self, imgs, attn_splits_list: Optional[Sequence[int]] = None, keep_raw_feats: bool = False,
                wo_self_attn: bool = False, pair_ids: Optional[Sequence[int]] = None, **kwargs):
        """``pair_ids`` (not in the reference): encode only these view pairs (indices into the (0,1), (0,2), (1,2) list) -- the
        multi-GPU encoder split of matchnerf_b200.sharding; the returned tensors then hold len(pair_ids) pairs."""
        with _matmul_precision(self.matmul_precision):
            return self._forward(imgs, attn_splits_list, keep_raw_feats, wo_self_attn, pair_ids)

    def _position(self, h, w, splits, device):
        pkey = (h, w, splits, device)
        if getattr(self, "_pos_cache", None) is None or self._pos_cache[0] != pkey:
            pos = sine_position(h // splits, w // splits, self.feature_channels, device).repeat(1, splits, splits)
            self._pos_cache = (pkey, pos, pos.permute(1, 2, 0).reshape(h * w, self.feature_channels).contiguous())
        return self._pos_cache[1]

    # Inference on the GPU keeps the activations in ONE layout from the backbone's last convolution to the up-sampler's first:
    # channels-last == tokens [image, h*w, C].  The reference-shaped path around the transformer (stack the pairs in NCHW, add the
    # position, cat + transpose to tokens, per-block cat of the swapped halves, transpose back, cat + convert for the up-sampler)
    # was ~25 copy / add kernels per encoder call; here: one add (position, V images), one gather (pairs), views.
    token_path = os.environ.get("MNF_TOKEN_PATH", "1") != "0"

    def _token_path_ok(self, imgs) -> bool:
        return (self.token_path and imgs.is_cuda and imgs.dtype == torch.float32 and not torch.is_grad_enabled()
                and self.backbone.fast_dtype is not None and TransformerLayer.fused_proj and TransformerLayer.fused_block
                and TransformerLayer.ffn_dtype == torch.float16 and self.feature_channels == 128)

    def _forward_tokens(self, imgs, B, V, pairs, splits, keep_raw_feats, wo_self_attn):
        C = self.feature_channels
        base = self.backbone._forward_nhwc(self.normalize_images(imgs).reshape(B * V, 3, *imgs.shape[-2:]), keep_nhwc=True)
        h, w = base.shape[-2:]
        if h % splits or w % splits:
            raise ValueError(f"feature map {h}x{w} not divisible by attn_splits {splits}")
        tok = base.permute(0, 2, 3, 1).reshape(B * V, h * w, C)              # a view of the channels-last convolution output
        self._position(h, w, splits, base.device)
        tok = tok + self._pos_cache[2]
        P = len(pairs)
        ikey = (B, V, tuple(pairs), base.device)
        if getattr(self, "_pair_index", None) is None or self._pair_index[0] != ikey:
            idx = [b * V + a for b in range(B) for a, _ in pairs] + [b * V + c for b in range(B) for _, c in pairs]
            self._pair_index = (ikey, torch.tensor(idx, dtype=torch.long, device=base.device))
        x = tok.index_select(0, self._pair_index[1])                         # [2BP, hw, C]: feature0 of every pair, then feature1
        x = self.transformer.forward_tokens(x, B * P, h, w, splits, wo_self_attn)
        xv = x.view(2 * B * P, h, w, C).permute(0, 3, 1, 2)                  # [2BP, C, h, w] with channels-last strides (no copy)
        out0, out1 = [], []
        if keep_raw_feats:
            raw = xv.contiguous()                                            # NCHW, as the reference-shaped path returns them
            out0.append(raw[: B * P].reshape(B, P, C, h, w))
            out1.append(raw[B * P:].reshape(B, P, C, h, w))
        if self.feature_upsampler == "network":
            up = self.featup_net(xv.contiguous(memory_format=self.upsampler_memory_format))   # already channels-last: no copy
            f0, f1 = up[: B * P], up[B * P:]
        else:
            f0, f1 = xv[: B * P], xv[B * P:]
        out0.append(f0.reshape(B, P, *f0.shape[1:]))
        out1.append(f1.reshape(B, P, *f1.shape[1:]))
        return {"aug_feat0s": out0, "aug_feat1s": out1}

    def _forward(self, imgs, attn_splits_list, keep_raw_feats, wo_self_attn, pair_ids=None):
        B, V, _, H, W = imgs.shape
        if H == 756 and W == 1008:     # IBRNet setting: pad to a size divisible by 16 (gmflow.py:99-103)
            imgs = F.interpolate(imgs.reshape(B * V, 3, H, W), size=(768, 1024), mode="bilinear",
                                 align_corners=True).reshape(B, V, 3, 768, 1024)
        splits = int((attn_splits_list or [2])[0])
        pairs = [(a, b) for a in range(V - 1) for b in range(a + 1, V)]
        if pair_ids is not None:
            pairs = [pairs[i] for i in pair_ids]
        if self._token_path_ok(imgs):
            return self._forward_tokens(imgs, B, V, pairs, splits, keep_raw_feats, wo_self_attn)
        base = self.backbone(self.normalize_images(imgs).reshape(B * V, 3, *imgs.shape[-2:])
                             .contiguous(memory_format=self.backbone_memory_format))
        base = base.reshape(B, V, *base.shape[1:])
        f0 = torch.stack([base[:, a] for a, _ in pairs], 1).flatten(0, 1)        # [B*P, C, h, w]
        f1 = torch.stack([base[:, b] for _, b in pairs], 1).flatten(0, 1)
        h, w = f0.shape[-2:]
        if h % splits or w % splits:
            raise ValueError(f"feature map {h}x{w} not divisible by attn_splits {splits}")
        pos = self._position(h, w, splits, f0.device)
        f0, f1 = self.transformer(f0 + pos, f1 + pos, splits, wo_self_attn)
        P = len(pairs)
        out0, out1 = [], []
        if keep_raw_feats:
            out0.append(f0.reshape(B, P, *f0.shape[1:]))
            out1.append(f1.reshape(B, P, *f1.shape[1:]))
        if self.feature_upsampler == "network":
            up = self.featup_net(torch.cat([f0, f1], 0).contiguous(memory_format=self.upsampler_memory_format))
            f0, f1 = up[: B * P], up[B * P:]
        out0.append(f0.reshape(B, P, *f0.shape[1:]))
        out1.append(f1.reshape(B, P, *f1.shape[1:]))
        return {"aug_feat0s": out0, "aug_feat1s": out1}
