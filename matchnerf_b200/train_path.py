"""Training-time render path (BASELINE configs[4]: ``configs/train.yaml``; reference ``coach.py:215-243`` ->
``models/matchnerf.py:51-55`` random rays -> ``:88-143`` render -> ``loss.backward()``).

What runs where in a training step
  * K-gather forward AND backward: this repo's CUDA kernels (``mnf_gather_cossim_fwd`` / ``mnf_gather_cossim_bwd``) behind one
    ``torch.autograd.Function`` -- the six ``F.grid_sample`` calls and the 30 grouped cosine similarities per sample of
    ``query_cond_info`` and their scatter-add backward, the part of the step that is not GEMM-shaped;
  * the conditional MLP, the ray transformer and the compositing of the 131,072 samples of a step (1,024 rays x 128): dense
    linear algebra on flat sample lists, evaluated with library GEMMs (cuBLAS) and differentiated by autograd -- the fused
    tcgen05 inference kernel keeps no activations, and a hand-written backward of it is the next step;
  * the encoder: its layers run the PyTorch ops the reference uses when gradients are requested (``gmflow.py``), the
    split-window attention as index-gathered batched GEMMs (``window_attention_autograd``).

The functions below are written on flat [N = R*S, .] sample lists (not the reference's [B, R, S, .] tensors) and take the decoder
parameters from ``matchnerf_b200.cond_nerf.CondNeRF`` (same state_dict keys as the reference).
"""
from __future__ import annotations

import math
import os
from typing import Tuple

import torch
import torch.nn.functional as F

from . import capi
from .utils import get_opt


class GatherCondFn(torch.autograd.Function):
    """cond [R*S, 22] = query_cond_info(feature maps) for the rays ``ray_idx`` of one batch item; backward scatters the gradient of
    the ten cosine similarities into the two feature maps (the colours / masks and the sample positions carry no gradient)."""

    @staticmethod
    def forward(fctx, feat8, feat4, lib_ctx, packed, sc, ray_idx, jitter, S):
        # ``packed`` = the scene packed from feat8 / feat4 (fp16 channels-last, what the kernels read), ``sc`` its camera block
        cond, _ = lib_ctx.gather_cossim(sc, S, ray_idx=ray_idx, jitter=jitter)
        fctx.saved = (lib_ctx, packed, sc, ray_idx, jitter, S)
        return cond

    @staticmethod
    def backward(fctx, dcond):
        lib_ctx, packed, sc, ray_idx, jitter, S = fctx.saved
        g8, g4 = lib_ctx.gather_cossim_bwd(sc, S, dcond.contiguous().float(), ray_idx=ray_idx, jitter=jitter)
        return (g8, g4) + (None,) * 6


def sample_geometry(sc: "capi.Scene", ray_idx: torch.Tensor, jitter, S: int, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(ndc [R,S,3] of the samples in source view 0, unit view direction in that view's frame [R,3], depth [R,S]) from the camera
    block the kernels use (the mnf_scene struct: float64-inverted target pose).  misc/camera.py:255-286, :351-379;
    models/matchnerf.py:129-134, :163-181 (legacy coordinates).  No gradient flows through the geometry."""
    f = lambda a, shape: torch.tensor(list(a), dtype=torch.float32, device=device).view(*shape)
    c2w, Kinv = f(sc.tgt_c2w, (3, 4)), f(sc.tgt_Kinv, (3, 3))
    near, far = float(sc.tgt_near_far[0]), float(sc.tgt_near_far[1])
    E0, K0 = f(sc.src_w2c[0], (3, 4)), f(sc.src_K[0], (3, 3))
    n0, f0 = float(sc.src_near_far[0][0]), float(sc.src_near_far[0][1])
    W, H = int(sc.W), int(sc.H)
    idx = ray_idx.to(device)
    pix = torch.stack([(idx % W).float(), torch.div(idx, W, rounding_mode="floor").float(), torch.ones_like(idx, dtype=torch.float32)], -1)
    d = (pix @ Kinv.T) @ c2w[:, :3].T                                        # [R,3] (the translation cancels in p - o)
    o = c2w[:, 3]
    steps = torch.arange(S, device=device, dtype=torch.float32)[None, :]
    u = jitter.to(device).view(-1, S) if jitter is not None else 0.0
    depth = near + (steps + u) / (S - 1) * (far - near)                      # [R,S]
    pts = o[None, None, :] + depth[..., None] * d[:, None, :]                # [R,S,3]
    q = (pts @ E0[:, :3].T + E0[:, 3]) @ K0.T
    ndc = torch.stack([q[..., 0] / q[..., 2] / (W - 1), q[..., 1] / q[..., 2] / (H - 1), (q[..., 2] - n0) / (f0 - n0)], -1)
    dirs = F.normalize(d, dim=-1) @ E0[:, :3].T
    return ndc, dirs, depth


def _posenc(x: torch.Tensor, L: int) -> torch.Tensor:
    """[x, sin(2^k x) (k-major), cos(2^k x) (k-major)]: cond_nerf.py:108-116 (legacy: no pi) + :56-57."""
    freq = 2.0 ** torch.arange(L, dtype=torch.float32, device=x.device)
    spec = (x[..., None, :] * freq[:, None]).flatten(-2)
    return torch.cat([x, spec.sin(), spec.cos()], -1)


def decode_samples(dec, opt, ndc: torch.Tensor, dirs: torch.Tensor, cond: torch.Tensor, S: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """CondNeRF.forward (models/rfdecoder/cond_nerf.py:52-100) on a flat sample list: ndc [R,S,3], dirs [R,3], cond [R*S,22]
    -> (rgb [R,S,3], sigma [R,S]).  Differentiable in the decoder parameters and in ``cond``."""
    R = ndc.shape[0]
    enc = _posenc(ndc.reshape(R * S, 3), int(get_opt(opt, "decoder.posenc.L_3D", 10)))
    gate = dec.pts_bias(cond)
    skips = list(get_opt(opt, "decoder.skip", [4]))
    h = enc
    for i, lin in enumerate(dec.pts_linears):
        h = F.relu(lin(h) * gate)
        if i in skips:
            h = torch.cat([enc, h], -1)
    x = dec.alpha_linear(h).view(R, S, 16)
    if bool(get_opt(opt, "decoder.raytrans_posenc", False)):
        pos = torch.arange(S, dtype=torch.float64)[:, None] / torch.pow(torch.tensor(10000.0, dtype=torch.float64),
                                                                         2.0 * (torch.arange(16) // 2).double() / 16.0)[None, :]
        tab = torch.where((torch.arange(16) % 2 == 0)[None, :], pos.sin(), pos.cos()).float().to(x.device)
        x = x + tab[None]
    seen = cond[:, 19:22].sum(-1).view(R, S)                                  # views that see the sample
    ra = dec.ray_attention
    q = ra.w_qs(x).view(R, S, 4, 4).transpose(1, 2)                           # [R, head, S, 4]
    k = ra.w_ks(x).view(R, S, 4, 4).transpose(1, 2)
    v = ra.w_vs(x).view(R, S, 4, 4).transpose(1, 2)
    att = (q * 0.5) @ k.transpose(-1, -2)                                     # temperature sqrt(d_k) = 2
    # cond_nerf.py:83 + ray_transformer.py:19: the mask broadcasts over the KEY axis, i.e. it disables whole QUERY rows
    att = att.masked_fill((seen <= 1)[:, None, :, None], -1e9)
    y = (att.softmax(-1) @ v).transpose(1, 2).reshape(R, S, 16)
    y = ra.layer_norm(ra.fc(y) + x)
    sigma = dec.out_alpha_linear(y).view(R, S)
    if bool(get_opt(opt, "decoder.density_maskfill", False)):
        sigma = sigma.masked_fill(seen < 1, 0.0)
    feat = dec.feature_linear(h)
    hv = F.relu(dec.views_linears[0](torch.cat([feat, dirs[:, None, :].expand(R, S, 3).reshape(R * S, 3)], -1)))
    rgb = torch.sigmoid(dec.rgb_linear(hv)).view(R, S, 3)
    return rgb, sigma


def composite_samples(rgb: torch.Tensor, sigma: torch.Tensor, depth: torch.Tensor, setbg_opaque: bool):
    """NeRF.composite with wo_render_interval (models/rfdecoder/nerf.py:101-124): alpha = 1 - exp(-sigma),
    T_i = exp(-sum_{j<i} sigma_j).  rgb [R,S,3], sigma / depth [R,S] -> rgb [R,3], depth [R,1], opacity [R,1]."""
    alpha = 1.0 - torch.exp(-sigma)
    T = torch.exp(-(torch.cumsum(sigma, -1) - sigma))
    w = T * alpha
    out_rgb = (w[..., None] * rgb).sum(1)
    opacity = w.sum(1, keepdim=True)
    out_depth = (w * depth).sum(1, keepdim=True)
    if setbg_opaque:
        out_rgb = out_rgb + (1.0 - opacity)
    return out_rgb, out_depth, opacity


def render_rays_train(model, opt, tgt_pose, ray_idx, ref_poses, ref_images, ref_feats_list, b: int, stratified: bool):
    """One batch item's training render: (rgb [R,3], depth [R,1], opacity [R,1]) with gradients to the encoder (through the
    feature maps) and to the decoder parameters."""
    dec = model._unwrap(model.nerf_dec)
    dev = ref_images.device
    if getattr(model, "local_radius", 0) > 0:
        raise NotImplementedError("feature_sample_local_radius > 0 has no backward kernel (forward / inference only)")
    lib_ctx = capi.get_context(dev)
    S = int(get_opt(opt, "nerf.sample_intvs", 128))
    R = int(ray_idx.numel())
    jitter = torch.rand((R, S), device=dev) if stratified else None        # matchnerf.py:168-169: u ~ U[0,1) per sample
    cpu = lambda t: t.detach().float().cpu()
    feat8, feat4 = ref_feats_list[0][b], ref_feats_list[1][b]
    packed = lib_ctx.pack_scene([feat8.detach().float(), feat4.detach().float()], ref_images[b].detach().float(),
                                cpu(ref_poses["extrinsics"][b]), cpu(ref_poses["intrinsics"][b]), cpu(ref_poses["near_fars"][b]))
    sc = packed.c_scene(cpu(tgt_pose["extrinsics"][b]), cpu(tgt_pose["intrinsics"][b]), cpu(tgt_pose["near_fars"][b]))
    cond = GatherCondFn.apply(feat8, feat4, lib_ctx, packed, sc, ray_idx, jitter, S)
    # the torch-side decoder inputs come from the SAME camera block the gather kernel read (float64-inverted target pose)
    ndc, dirs, depth = sample_geometry(sc, ray_idx, jitter, S, dev)
    rgb, sigma = decode_samples(dec, opt, ndc, dirs, cond, S)
    return composite_samples(rgb, sigma, depth, bool(model.nerf_setbg_opaque))


class training_precision:
    """Scoped cuBLAS / cuDNN math mode for a WHOLE training step -- forward AND backward: autograd launches the backward GEMMs after
    ``forward`` has returned, so the encoder's own scope (``GMFlow.matmul_precision``) does not cover them, and the decoder / ray
    transformer GEMMs of the training path are never inside it.  A torch.profiler pass of the step (tools/r02_train_prof.py) showed
    57 of 124 ms per two steps in fp32 ``simt_sgemm`` kernels.
      None    leave the process-wide flags alone (PyTorch's defaults = what the reference runs: fp32 GEMMs, TF32 convolutions);
      "tf32"  10-bit-mantissa operands, fp32 accumulation on the tensor cores for every GEMM / convolution of the step (the
              "bf16 / TF32 training step" of SURVEY 8f rank 2): 47 ms instead of 67-79 ms per step on B200 -- but OPT-IN: on the
              random-initialised test encoder the backward in TF32 moves the ENCODER gradient to cosine 0.69 against the fp32
              backward (decoder gradient: 0.997; tests/test_gpu_train.py) -- twelve attention layers amplify the operand rounding;
              whether trained weights behave better cannot be measured here (no checkpoint offline);
      "fp32"  fp32 everywhere, convolutions included."""

    def __init__(self, mode=None):
        if mode not in (None, "tf32", "fp32"):
            raise ValueError("precision must be None, 'tf32' or 'fp32'")
        self.mode = mode

    def __enter__(self):
        self.prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        if self.mode is not None:
            torch.backends.cuda.matmul.allow_tf32 = self.mode == "tf32"
            torch.backends.cudnn.allow_tf32 = self.mode == "tf32"
        return self

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.prev
        return False


# training-step attention through F.scaled_dot_product_attention (library) instead of explicit GEMMs: measured SLOWER on B200 at DTU
# size (fp32 inputs: the memory-efficient backend; 74 ms vs 66.5 ms per training step, tools/r02_train_ab.py) -- opt-in only
USE_SDPA = os.environ.get("MNF_TRAIN_SDPA", "0") != "0"


def window_attention_autograd(q, k, v, h: int, w: int, num_splits: int, with_shift: bool):
    """Split-window attention (models/gmflow/transformer.py:46-105, :19-43) as index-gathered batched GEMMs, for the training
    step only (inference uses the tcgen05 kernel behind mnf_window_attn_fwd).  Tokens are gathered per window through an index
    map that encodes the cyclic shift; keys of another shift region get -100 added, as the reference's mask does."""
    B, L, C = q.shape
    wh, ww = h // num_splits, w // num_splits
    sh, sw = (wh // 2, ww // 2) if (with_shift and num_splits > 1) else (0, 0)
    dev = q.device
    ys, xs = torch.arange(h, device=dev), torch.arange(w, device=dev)
    src_y, src_x = (ys + sh) % h, (xs + sw) % w                              # rolled position p holds token (p + shift) mod size
    tok = (src_y[:, None] * w + src_x[None, :])                               # [h, w] token id at each rolled position
    tok = tok.view(num_splits, wh, num_splits, ww).permute(0, 2, 1, 3).reshape(num_splits * num_splits, wh * ww)

    def region(n, win, shift):
        p = torch.arange(n, device=dev)
        return torch.where(p < n - win, 0, torch.where(p < n - shift, 1, 2)) if shift > 0 else torch.zeros(n, dtype=torch.long, device=dev)
    reg = (region(h, wh, sh)[:, None] * 3 + region(w, ww, sw)[None, :])
    reg = reg.view(num_splits, wh, num_splits, ww).permute(0, 2, 1, 3).reshape(num_splits * num_splits, wh * ww)
    bias = torch.where(reg[:, :, None] == reg[:, None, :], 0.0, -100.0).to(q.dtype)      # [windows, Lw, Lw]
    g = lambda t: t[:, tok]                                                   # [B, windows, Lw, C]
    if USE_SDPA and q.is_cuda:
        # library fused attention (windows in the head dimension, the shift mask as an additive bias): no [B, windows, Lw, Lw]
        # score / probability tensors saved for backward (157 MB each per call at DTU size)
        out_w = F.scaled_dot_product_attention(g(q), g(k), g(v), attn_mask=bias if (sh or sw) else None)
    else:
        att = (g(q) @ g(k).transpose(-1, -2)) * (1.0 / math.sqrt(C)) + bias[None]
        out_w = att.softmax(-1) @ g(v)
    flat = tok.reshape(-1)
    inv = torch.empty_like(flat)
    inv[flat] = torch.arange(flat.numel(), device=dev)                        # token id -> its row in the window-major list
    return out_w.reshape(B, -1, C)[:, inv]
