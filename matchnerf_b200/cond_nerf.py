"""Host-side mirror of the reference radiance-field decoder module (``models/rfdecoder/cond_nerf.py``).

It is a parameter container with the reference's module tree / state_dict keys (so ``nerf_dec`` checkpoints
load with ``strict=True``) plus the glue that hands the weights to the CUDA library.  The arithmetic lives
in ``csrc/decoder_*.cu`` behind ``mnf_decoder_composite_fwd`` / ``mnf_render_rays_fwd``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import capi
from .utils import get_opt


class _RayAttentionParams(nn.Module):
    """Parameters of the 4-head ray transformer (models/rfdecoder/ray_transformer.py:29-47)."""

    def __init__(self, n_head=4, d_model=16, d_k=4, d_v=4):
        super().__init__()
        self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(d_model, n_head * d_v, bias=False)
        self.fc = nn.Linear(n_head * d_v, d_model, bias=False)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)


class CondNeRF(nn.Module):
    """Parameter tree of models/rfdecoder/cond_nerf.py:15-50 (view-dependent branch)."""

    def __init__(self, opt):
        super().__init__()
        W = int(get_opt(opt, "decoder.net_width", 128))
        D = int(get_opt(opt, "decoder.net_depth", 6))
        skips = list(get_opt(opt, "decoder.skip", [4]))
        L3 = int(get_opt(opt, "decoder.posenc.L_3D", 10))
        Lv = int(get_opt(opt, "decoder.posenc.L_view", 0))
        groups = get_opt(opt, "encoder.cos_n_group", [2, 8])
        groups = [groups] if isinstance(groups, int) else list(groups)
        n_views = int(get_opt(opt, "n_src_views", 3))
        if (W, D, skips, L3, Lv, groups, n_views) != (128, 6, [4], 10, 0, [2, 8], 3) or not get_opt(opt, "nerf.view_dep", True):
            raise NotImplementedError(
                "matchnerf_b200 builds the decoder architecture every shipped reference config uses "
                "(net_width 128, net_depth 6, skip [4], L_3D 10, L_view 0, cos_n_group [2, 8], 3 views, view_dep); "
                f"got width={W} depth={D} skip={skips} L_3D={L3} L_view={Lv} groups={groups} views={n_views}")
        in3d = 3 + 6 * L3
        in_feat = sum(groups) + n_views * 4
        act = getattr(nn, str(get_opt(opt, "decoder.raytrans_act", "ReLU")))
        self.pts_linears = nn.ModuleList([nn.Linear(in3d, W)] + [nn.Linear(W + in3d if i in skips else W, W) for i in range(D - 1)])
        self.pts_bias = nn.Linear(in_feat, W)
        self.views_linears = nn.ModuleList([nn.Linear(3 + W, W // 2)])
        self.alpha_linear = nn.Sequential(nn.Linear(W, 16), act())
        self.ray_attention = _RayAttentionParams()
        self.out_alpha_linear = nn.Sequential(nn.Linear(16, 16), act(), nn.Linear(16, 1), nn.ReLU())
        self.feature_linear = nn.Linear(W, W)
        self.rgb_linear = nn.Linear(W // 2, 3)
        for grp in (self.pts_linears, self.views_linears, self.feature_linear, self.alpha_linear, self.rgb_linear):
            for m in grp.modules():                       # cond_nerf.py:46-50, :102-106
                if isinstance(m, nn.Linear):
                    nn.init.kaiming_normal_(m.weight)
                    nn.init.zeros_(m.bias)

    # ---- weights -> library
    def _version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def sync_to_library(self, ctx: Optional["capi.Context"] = None) -> "capi.Context":
        """(Re)upload the weights into the library context when they changed (load_state_dict, optimiser step)."""
        dev = next(self.parameters()).device
        ctx = ctx or capi.get_context(dev)
        owner = (id(self), self._version())
        if ctx.decoder_owner != owner or not ctx.decoder_loaded:      # someone else's (or stale) weights are resident
            ctx.load_decoder(self.state_dict(), owner=owner)
        return ctx

    def decoder_cfg(self, opt) -> "capi.DecoderCfg":
        cfg = capi.DecoderCfg()
        cfg.n_samples = int(get_opt(opt, "nerf.sample_intvs", 128))
        cfg.raytrans_act = {"ReLU": 0, "ELU": 1}[str(get_opt(opt, "decoder.raytrans_act", "ReLU"))]
        cfg.raytrans_posenc = int(bool(get_opt(opt, "decoder.raytrans_posenc", False)))
        cfg.density_maskfill = int(bool(get_opt(opt, "decoder.density_maskfill", False)))
        return cfg

    def forward(self, opt, points_3D, ray_unit=None, cond_info=None, mode=None):
        """models/rfdecoder/cond_nerf.py:52-100 on explicit tensors: points_3D [B,R,S,3] (view-0 NDC), ray_unit [B,R,S,3],
        cond_info = dict(feat_info [B,R,S,10], color_info [B,R,S,9], mask_info [B,R,S,3]) -> (rgb [B,R,S,3], alpha [B,R,S]).
        Runs the fp32 CUDA kernel (mnf_decoder_samples_fwd); MatchNeRF.render uses the fused tcgen05 kernel instead."""
        if not points_3D.is_cuda:
            raise RuntimeError("matchnerf_b200: the decoder only exists as CUDA kernels (no CPU path)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("matchnerf_b200: the backward kernels are not built yet; call under torch.no_grad()")
        if ray_unit is None or cond_info is None:
            raise ValueError("CondNeRF.forward needs ray_unit (nerf.view_dep) and cond_info")
        ctx = self.sync_to_library()
        cfg = self.decoder_cfg(opt)
        B, R, S, _ = points_3D.shape
        cfg.n_samples = S
        cond = torch.cat([cond_info["feat_info"], cond_info["color_info"], cond_info["mask_info"]], dim=-1).float()
        out = ctx.decoder_samples(cfg, points_3D.float().reshape(B * R, S, 3), ray_unit.float().expand(B, R, S, 3).reshape(B * R, S, 3),
                                  cond.reshape(B * R * S, -1))
        out = out.view(B, R, S, 4)
        return out[..., :3], out[..., 3]

    @staticmethod
    def composite(opt, ray, rgb_samples, density_samples, depth_samples, setbg_opaque):
        """Alpha compositing on explicit sample tensors (models/rfdecoder/nerf.py:101-124) through mnf_composite_fwd:
        rgb_samples [B,R,S,3], density_samples [B,R,S], depth_samples [B,R,S,1] -> rgb [B,R,3], depth [B,R,1],
        opacity [B,R,1], prob [B,R,S,1].  (MatchNeRF.render composites inside the fused kernel.)"""
        if not get_opt(opt, "nerf.wo_render_interval", True):
            raise NotImplementedError("only wo_render_interval=True (all shipped configs) is built")
        if not rgb_samples.is_cuda:
            raise RuntimeError("matchnerf_b200: compositing only exists as a CUDA kernel (no CPU path)")
        B, R, S, _ = rgb_samples.shape
        ctx = capi.get_context(rgb_samples.device)
        rgb, depth, opac, prob = ctx.composite(rgb_samples.float().reshape(B * R, S, 3), density_samples.float().reshape(B * R, S),
                                               depth_samples.float().reshape(B * R, S), bool(setbg_opaque))
        return rgb.view(B, R, 3), depth.view(B, R, 1), opac.view(B, R, 1), prob.view(B, R, S, 1)
