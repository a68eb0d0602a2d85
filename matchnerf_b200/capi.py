"""ctypes binding of libmatchnerf_b200.so (the C ABI in include/matchnerf_b200.h).

PyTorch is used here only as the owner of device memory and streams: every call passes raw device
pointers plus the current CUDA stream.  There is no fallback: if the shared library is missing or a
call fails, a ``RuntimeError`` carrying ``mnf_last_error()`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MNF_LIB_PATH") or os.path.join(_HERE, "libmatchnerf_b200.so")   # override: A/B builds (tools/)

ABI_VERSION = 9
COND_DIM = 22
COND_PAD = 32
FEAT_CH = 256

# reference ``nerf_dec`` state_dict order (models/rfdecoder/cond_nerf.py:15-50); must match csrc/decoder_weights.cuh
DECODER_PARAM_ORDER = (
    [f"pts_linears.{i}.{n}" for i in range(6) for n in ("weight", "bias")]
    + ["pts_bias.weight", "pts_bias.bias", "views_linears.0.weight", "views_linears.0.bias",
       "alpha_linear.0.weight", "alpha_linear.0.bias",
       "ray_attention.w_qs.weight", "ray_attention.w_ks.weight", "ray_attention.w_vs.weight", "ray_attention.fc.weight",
       "ray_attention.layer_norm.weight", "ray_attention.layer_norm.bias",
       "out_alpha_linear.0.weight", "out_alpha_linear.0.bias", "out_alpha_linear.2.weight", "out_alpha_linear.2.bias",
       "feature_linear.weight", "feature_linear.bias", "rgb_linear.weight", "rgb_linear.bias"]
)


class DecoderCfg(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("raytrans_act", C.c_int32), ("raytrans_posenc", C.c_int32),
                ("density_maskfill", C.c_int32)]


class Scene(C.Structure):
    _fields_ = [("n_views", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("h0", C.c_int32), ("w0", C.c_int32), ("h1", C.c_int32), ("w1", C.c_int32),
                ("feat0", C.c_void_p), ("feat1", C.c_void_p), ("images", C.c_void_p),
                ("src_w2c", (C.c_float * 12) * 3), ("src_K", (C.c_float * 9) * 3), ("src_near_far", (C.c_float * 2) * 3),
                ("tgt_c2w", C.c_float * 12), ("tgt_Kinv", C.c_float * 9), ("tgt_near_far", C.c_float * 2),
                ("sample_local_radius", C.c_int32), ("sample_local_dilation", C.c_int32)]


class Rays(C.Structure):
    _fields_ = [("n_rays", C.c_int64), ("ray_idx", C.c_void_p), ("first_ray", C.c_int64), ("jitter", C.c_void_p)]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (built by ``matchnerf_b200.build``); fail loudly if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m matchnerf_b200.build` "
                           "(there is no CPU or PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, fp = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    lib.mnf_abi_version.restype = i32
    lib.mnf_last_error.restype = C.c_char_p
    lib.mnf_decoder_param_count.restype = i64
    lib.mnf_ctx_create.argtypes = [i32, C.POINTER(vp)]
    lib.mnf_ctx_destroy.argtypes = [vp]
    lib.mnf_decoder_load_host.argtypes = [vp, fp, i64]
    lib.mnf_packed_feature_halves.argtypes = [i32, i32, i32]
    lib.mnf_packed_feature_halves.restype = i64
    lib.mnf_pack_features.argtypes = [vp, fp, i32, i32, i32, vp, vp]
    lib.mnf_pack_images.argtypes = [vp, fp, i32, i32, i32, vp, vp]
    lib.mnf_gather_cossim_fwd.argtypes = [vp, C.POINTER(Scene), C.POINTER(Rays), i32, fp, vp, vp]
    lib.mnf_decoder_composite_fwd.argtypes = [vp, C.POINTER(Scene), C.POINTER(Rays), C.POINTER(DecoderCfg), fp, vp, i32,
                                              fp, fp, fp, fp, i32, vp]
    lib.mnf_query_cond_points_fwd.argtypes = [vp, C.POINTER(Scene), fp, i64, i32, fp, vp, vp]
    lib.mnf_decoder_samples_fwd.argtypes = [vp, C.POINTER(DecoderCfg), fp, fp, fp, i64, fp, vp]
    lib.mnf_composite_fwd.argtypes = [vp, fp, fp, fp, i64, i32, i32, fp, fp, fp, fp, vp]
    lib.mnf_instance_norm_fwd.argtypes = [vp, fp, fp, fp, i64, i32, i32, C.c_float, vp]
    lib.mnf_instance_norm_nhwc_fwd.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, C.c_float, vp]
    lib.mnf_gather_cossim_bwd.argtypes = [vp, C.POINTER(Scene), C.POINTER(Rays), i32, fp, fp, fp, vp]
    lib.mnf_token_layernorm_fwd.argtypes = [vp, vp, i32, fp, fp, C.c_float, fp, fp, fp, vp, i64, i32, vp]
    lib.mnf_token_block_weight_bytes.argtypes = [i32]
    lib.mnf_token_block_weight_bytes.restype = i64
    lib.mnf_token_block_pack_weights.argtypes = [vp, fp, fp, fp, fp, fp, fp, fp, vp, vp]
    lib.mnf_token_block_fwd.argtypes = [vp, fp, fp, vp, i32, C.c_float, fp, i64, i32, vp]
    lib.mnf_render_workspace_bytes.argtypes = [i64, i32, i32]
    lib.mnf_render_workspace_bytes.restype = i64
    lib.mnf_render_rays_fwd.argtypes = [vp, C.POINTER(Scene), C.POINTER(Rays), C.POINTER(DecoderCfg), i32, fp, fp, fp,
                                        vp, i64, i32, vp]
    lib.mnf_window_attn_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.mnf_window_attn_workspace_bytes.restype = i64
    lib.mnf_window_attn_fwd.argtypes = [vp, fp, fp, fp, fp, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp]
    lib.mnf_window_attn_proj_weight_bytes.restype = i64
    lib.mnf_window_attn_pack_proj_weights.argtypes = [vp, fp, fp, fp, vp, vp]
    lib.mnf_window_attn_proj_fwd.argtypes = [vp, fp, fp, vp, fp, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp]
    lib.mnf_image_metrics_fwd.argtypes = [vp, fp, fp, vp, i32, i32, i32, i32, i32, i32, C.c_float, vp, vp]
    lib.mnf_selftest_umma.argtypes = [vp, vp, fp, i32, i32, i32, vp]
    for name in ("mnf_ctx_create", "mnf_ctx_destroy", "mnf_decoder_load_host", "mnf_pack_features", "mnf_pack_images",
                 "mnf_gather_cossim_fwd", "mnf_decoder_composite_fwd", "mnf_render_rays_fwd", "mnf_window_attn_fwd",
                 "mnf_selftest_umma", "mnf_query_cond_points_fwd", "mnf_decoder_samples_fwd", "mnf_composite_fwd",
                 "mnf_instance_norm_fwd", "mnf_token_layernorm_fwd", "mnf_gather_cossim_bwd",
                 "mnf_instance_norm_nhwc_fwd", "mnf_token_block_pack_weights", "mnf_token_block_fwd",
                 "mnf_window_attn_pack_proj_weights", "mnf_window_attn_proj_fwd", "mnf_image_metrics_fwd"):
        getattr(lib, name).restype = i32
    if lib.mnf_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libmatchnerf_b200.so ABI {lib.mnf_abi_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mnf_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed with status {rc}: {msg}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _dev_f32(t: torch.Tensor, device: torch.device, name: str) -> torch.Tensor:
    if t.device != device:
        raise ValueError(f"{name} is on {t.device}, the context is bound to {device}")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def flatten_decoder_state(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """``nerf_dec`` state_dict -> the flat fp32 host blob mnf_decoder_load_host expects."""
    missing = [k for k in DECODER_PARAM_ORDER if k not in sd]
    if missing:
        raise KeyError(f"decoder state_dict lacks {missing}")
    return torch.cat([sd[k].detach().to("cpu", torch.float32).reshape(-1) for k in DECODER_PARAM_ORDER]).contiguous()


class PackedScene:
    """Device-resident packed feature maps / images of one encoded source-view triplet plus its cameras."""

    def __init__(self, feat0, feat1, images, H, W, h0, w0, h1, w1, src_w2c, src_K, src_nf, local_radius=0, local_dilation=1):
        self.feat0, self.feat1, self.images = feat0, feat1, images
        self.local_radius, self.local_dilation = int(local_radius), int(local_dilation)   # encoder.feature_sample_local_*
        self.H, self.W, self.h0, self.w0, self.h1, self.w1 = H, W, h0, w0, h1, w1
        self.src_w2c, self.src_K, self.src_nf = src_w2c, src_K, src_nf   # CPU float32 [3,3,4], [3,3,3], [3,2]

    def c_scene(self, tgt_w2c: torch.Tensor, tgt_K: torch.Tensor, tgt_nf: torch.Tensor) -> Scene:
        """Fill the C struct for one target camera (w2c [3,4], K [3,3], near_far [2]; any device)."""
        sc = Scene()
        sc.n_views, sc.H, sc.W = 3, self.H, self.W
        sc.h0, sc.w0, sc.h1, sc.w1 = self.h0, self.w0, self.h1, self.w1
        sc.feat0, sc.feat1, sc.images = self.feat0.data_ptr(), self.feat1.data_ptr(), self.images.data_ptr()
        w2c, K, nf = self.src_w2c.tolist(), self.src_K.tolist(), self.src_nf.tolist()
        for v in range(3):
            sc.src_w2c[v][:] = [x for row in w2c[v] for x in row]
            sc.src_K[v][:] = [x for row in K[v] for x in row]
            sc.src_near_far[v][:] = nf[v]
        # float64 inverse of the target pose, cast to fp32 (misc/camera.py:231-240); fp32 inverse of K (camera.py:221)
        sq = torch.eye(4, dtype=torch.float64)
        sq[:3, :] = tgt_w2c.detach().to("cpu", torch.float64)
        c2w = torch.linalg.inv(sq)[:3, :].to(torch.float32)
        kinv = torch.linalg.inv(tgt_K.detach().to("cpu", torch.float32))
        sc.tgt_c2w[:] = c2w.reshape(-1).tolist()
        sc.tgt_Kinv[:] = kinv.reshape(-1).tolist()
        sc.tgt_near_far[:] = tgt_nf.detach().to("cpu", torch.float32).tolist()
        sc.sample_local_radius, sc.sample_local_dilation = self.local_radius, self.local_dilation
        sc._owner = self        # the struct holds raw device pointers: keep the packed buffers alive as long as it is
        return sc


class Context:
    """One mnf_ctx: bound to a CUDA device, owns the packed decoder weights."""

    def __init__(self, device=None):
        self.lib = load()
        if not torch.cuda.is_available():
            raise RuntimeError("matchnerf_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        h = C.c_void_p()
        _check(self.lib.mnf_ctx_create(self.device.index, C.byref(h)), "mnf_ctx_create")
        self._h = h
        self.decoder_loaded = False
        # who loaded the weights currently packed in this context: (id(module), parameter version) for a CondNeRF that
        # synchronised itself, None after a direct load_decoder(state_dict).  A context is shared by every model on its
        # device, so a module must reload when ANOTHER owner's weights are resident (two checkpoints, EMA + live model, ...).
        self.decoder_owner = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mnf_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights
    def load_decoder(self, state_dict: Dict[str, torch.Tensor], owner=None) -> None:
        blob = flatten_decoder_state(state_dict)
        if blob.numel() != self.lib.mnf_decoder_param_count():
            raise ValueError(f"decoder has {blob.numel()} parameters, library expects {self.lib.mnf_decoder_param_count()}")
        with torch.cuda.device(self.device):
            _check(self.lib.mnf_decoder_load_host(self._h, blob.data_ptr(), blob.numel()), "mnf_decoder_load_host")
        self.decoder_loaded = True
        self.decoder_owner = owner

    # ---- packing
    def pack_scene(self, feats: Sequence[torch.Tensor], images: torch.Tensor, src_w2c: torch.Tensor, src_K: torch.Tensor,
                   src_nf: torch.Tensor, local_radius: int = 0, local_dilation: int = 1) -> PackedScene:
        """feats: [feat_coarse [V,256,h0,w0], feat_fine [V,256,h1,w1]] fp32 NCHW on device (get_img_feat layout
        without the batch dim); images [V,3,H,W] fp32 in [0,1]; cameras [V,3,4], [V,3,3], [V,2]."""
        f0 = _dev_f32(feats[0], self.device, "feat0")
        f1 = _dev_f32(feats[1], self.device, "feat1")
        im = _dev_f32(images, self.device, "images")
        V, ch, h0, w0 = f0.shape
        _, _, h1, w1 = f1.shape
        _, _, H, W = im.shape
        if V != 3 or ch != FEAT_CH or f1.shape[1] != FEAT_CH or im.shape[1] != 3:
            raise ValueError("expected 3 views x 256 channels feature maps and RGB images")
        # flat fp16 buffers in the library's packed layout (opaque to the caller; the size includes the tail padding)
        p0 = torch.empty(self.lib.mnf_packed_feature_halves(V, h0, w0), dtype=torch.float16, device=self.device)
        p1 = torch.empty(self.lib.mnf_packed_feature_halves(V, h1, w1), dtype=torch.float16, device=self.device)
        pi = torch.empty((V, H, W, 4), dtype=torch.float32, device=self.device)
        st = _stream(self.device)
        _check(self.lib.mnf_pack_features(self._h, f0.data_ptr(), V, h0, w0, p0.data_ptr(), st), "mnf_pack_features")
        _check(self.lib.mnf_pack_features(self._h, f1.data_ptr(), V, h1, w1, p1.data_ptr(), st), "mnf_pack_features")
        _check(self.lib.mnf_pack_images(self._h, im.data_ptr(), V, H, W, pi.data_ptr(), st), "mnf_pack_images")
        return PackedScene(p0, p1, pi, H, W, h0, w0, h1, w1,
                           src_w2c.detach().to("cpu", torch.float32)[:, :3, :4].contiguous(),
                           src_K.detach().to("cpu", torch.float32).contiguous(),
                           src_nf.detach().to("cpu", torch.float32).contiguous(), local_radius, local_dilation)

    # ---- rays helper
    def _rays(self, n_rays: int, ray_idx: Optional[torch.Tensor], first_ray: int, jitter: Optional[torch.Tensor], S: int):
        r = Rays()
        keep = []
        if ray_idx is not None:
            ri = ray_idx.to(self.device, torch.int64).contiguous()
            keep.append(ri)
            n_rays = ri.numel()
            r.ray_idx = ri.data_ptr()
        else:
            r.ray_idx = None
        if jitter is not None:
            jt = _dev_f32(jitter, self.device, "jitter").reshape(n_rays, S)
            keep.append(jt)
            r.jitter = jt.data_ptr()
        else:
            r.jitter = None
        r.n_rays, r.first_ray = n_rays, first_ray
        return r, n_rays, keep

    # ---- kernels
    def gather_cossim(self, scene: Scene, S: int, ray_idx=None, first_ray=0, n_rays=0, jitter=None, want_f32=True,
                      want_f16=False):
        rays, R, keep = self._rays(n_rays, ray_idx, first_ray, jitter, S)
        c32 = torch.empty((R * S, COND_DIM), dtype=torch.float32, device=self.device) if want_f32 else None
        c16 = torch.empty((R * S, COND_PAD), dtype=torch.float16, device=self.device) if want_f16 else None
        _check(self.lib.mnf_gather_cossim_fwd(self._h, C.byref(scene), C.byref(rays), S, _ptr(c32), _ptr(c16),
                                              _stream(self.device)), "mnf_gather_cossim_fwd")
        return c32, c16

    def gather_cossim_bwd(self, scene: Scene, S: int, dcond: torch.Tensor, ray_idx=None, first_ray=0, n_rays=0, jitter=None):
        """d(loss)/d(cond) [R*S,22] -> gradients of the two feature maps, [V,256,h,w] fp32 each (NCHW, the layout the encoder
        produced them in); see mnf_gather_cossim_bwd.  Same scene / rays / jitter as the forward call."""
        rays, R, keep = self._rays(n_rays, ray_idx, first_ray, jitter, S)
        dc = _dev_f32(dcond, self.device, "dcond")
        if dc.shape != (R * S, COND_DIM):
            raise ValueError(f"dcond {tuple(dc.shape)} != {(R * S, COND_DIM)}")
        V = 3
        g0 = torch.zeros((V, scene.h0, scene.w0, FEAT_CH), dtype=torch.float32, device=self.device)
        g1 = torch.zeros((V, scene.h1, scene.w1, FEAT_CH), dtype=torch.float32, device=self.device)
        _check(self.lib.mnf_gather_cossim_bwd(self._h, C.byref(scene), C.byref(rays), S, dc.data_ptr(), g0.data_ptr(), g1.data_ptr(),
                                              _stream(self.device)), "mnf_gather_cossim_bwd")
        perm = self._unpack_perm()
        return (g0.index_select(3, perm).permute(0, 3, 1, 2).contiguous(), g1.index_select(3, perm).permute(0, 3, 1, 2).contiguous())

    def _unpack_perm(self) -> torch.Tensor:
        """perm[c] = packed position of channel c (mnf_pack_features: position 8 l + e <-> channel (e < 4 ? 0 : 128) + 4 l + (e & 3))."""
        if getattr(self, "_perm", None) is None:
            c = torch.arange(FEAT_CH)
            half, within = c // 128, c % 128
            self._perm = (8 * (within // 4) + 4 * half + within % 4).to(self.device)
        return self._perm

    def decoder_composite(self, scene: Scene, cfg: DecoderCfg, cond_f32=None, cond_f16=None, ray_idx=None, first_ray=0,
                          n_rays=0, jitter=None, setbg_opaque=False, impl=0, want_aux=False):
        S = cfg.n_samples
        rays, R, keep = self._rays(n_rays, ray_idx, first_ray, jitter, S)
        rgb = torch.empty((R, 3), dtype=torch.float32, device=self.device)
        depth = torch.empty((R,), dtype=torch.float32, device=self.device)
        opac = torch.empty((R,), dtype=torch.float32, device=self.device)
        aux = torch.empty((R * S, 4), dtype=torch.float32, device=self.device) if want_aux else None
        _check(self.lib.mnf_decoder_composite_fwd(self._h, C.byref(scene), C.byref(rays), C.byref(cfg), _ptr(cond_f32),
                                                  _ptr(cond_f16), int(setbg_opaque), rgb.data_ptr(), depth.data_ptr(),
                                                  opac.data_ptr(), _ptr(aux), impl, _stream(self.device)),
               "mnf_decoder_composite_fwd")
        return rgb, depth, opac, aux

    # ---- the reference's unfused per-sample methods on explicit tensors
    def query_cond_points(self, scene: Scene, points: torch.Tensor, want_f16=False):
        """points [R,S,3] world-space samples -> cond [R*S,22] fp32 (and [R*S,32] fp16): query_cond_info on explicit points."""
        pts = _dev_f32(points, self.device, "points")
        R, S = pts.shape[0], pts.shape[1]
        c32 = torch.empty((R * S, COND_DIM), dtype=torch.float32, device=self.device)
        c16 = torch.empty((R * S, COND_PAD), dtype=torch.float16, device=self.device) if want_f16 else None
        _check(self.lib.mnf_query_cond_points_fwd(self._h, C.byref(scene), pts.data_ptr(), R, S, c32.data_ptr(), _ptr(c16),
                                                  _stream(self.device)), "mnf_query_cond_points_fwd")
        return c32, c16

    def decoder_samples(self, cfg: DecoderCfg, pts_ndc: torch.Tensor, ray_unit: torch.Tensor, cond_f32: torch.Tensor):
        """CondNeRF.forward on explicit tensors: pts_ndc / ray_unit [R,S,3], cond [R*S,22] -> [R*S,4] (rgb, density)."""
        ndc = _dev_f32(pts_ndc, self.device, "pts_ndc")
        dirs = _dev_f32(ray_unit, self.device, "ray_unit")
        cond = _dev_f32(cond_f32, self.device, "cond")
        R, S = ndc.shape[0], ndc.shape[1]
        if S != cfg.n_samples or dirs.shape != ndc.shape or cond.numel() != R * S * COND_DIM:
            raise ValueError(f"decoder_samples: shapes {tuple(ndc.shape)} / {tuple(dirs.shape)} / {tuple(cond.shape)} do not match S={cfg.n_samples}")
        out = torch.empty((R * S, 4), dtype=torch.float32, device=self.device)
        _check(self.lib.mnf_decoder_samples_fwd(self._h, C.byref(cfg), ndc.data_ptr(), dirs.data_ptr(), cond.data_ptr(), R,
                                                out.data_ptr(), _stream(self.device)), "mnf_decoder_samples_fwd")
        return out

    def composite(self, rgb: torch.Tensor, sigma: torch.Tensor, depth: torch.Tensor, setbg_opaque=False, want_prob=True):
        """NeRF.composite: rgb [R,S,3], sigma [R,S], depth [R,S] -> rgb [R,3], depth [R], opacity [R], prob [R,S]."""
        c = _dev_f32(rgb, self.device, "rgb")
        sg = _dev_f32(sigma, self.device, "sigma")
        d = _dev_f32(depth, self.device, "depth")
        R, S = sg.shape
        if c.shape != (R, S, 3) or d.shape != (R, S):
            raise ValueError(f"composite: shapes {tuple(c.shape)} / {tuple(sg.shape)} / {tuple(d.shape)} do not match")
        o_rgb = torch.empty((R, 3), dtype=torch.float32, device=self.device)
        o_d = torch.empty((R,), dtype=torch.float32, device=self.device)
        o_o = torch.empty((R,), dtype=torch.float32, device=self.device)
        prob = torch.empty((R, S), dtype=torch.float32, device=self.device) if want_prob else None
        _check(self.lib.mnf_composite_fwd(self._h, c.data_ptr(), sg.data_ptr(), d.data_ptr(), R, S, int(setbg_opaque),
                                          o_rgb.data_ptr(), o_d.data_ptr(), o_o.data_ptr(), _ptr(prob), _stream(self.device)),
               "mnf_composite_fwd")
        return o_rgb, o_d, o_o, prob

    def render_rays(self, scene: Scene, cfg: DecoderCfg, ray_idx=None, first_ray=0, n_rays=0, jitter=None,
                    setbg_opaque=False, impl=0, out=None, workspace=None):
        """MatchNeRF.render for one slice.  Returns (rgb [R,3], depth [R], opacity [R])."""
        S = cfg.n_samples
        rays, R, keep = self._rays(n_rays, ray_idx, first_ray, jitter, S)
        if out is None:
            out = (torch.empty((R, 3), dtype=torch.float32, device=self.device),
                   torch.empty((R,), dtype=torch.float32, device=self.device),
                   torch.empty((R,), dtype=torch.float32, device=self.device))
        need = self.lib.mnf_render_workspace_bytes(R, S, impl)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty((max(need, 256),), dtype=torch.uint8, device=self.device)
        _check(self.lib.mnf_render_rays_fwd(self._h, C.byref(scene), C.byref(rays), C.byref(cfg), int(setbg_opaque),
                                            out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), workspace.data_ptr(),
                                            workspace.numel(), impl, _stream(self.device)), "mnf_render_rays_fwd")
        return out

    def instance_norm(self, x: torch.Tensor, mode: int = 1, residual: Optional[torch.Tensor] = None, eps: float = 1e-5):
        """Fused InstanceNorm2d(+ReLU)(+residual, ReLU) on a contiguous NCHW fp32 tensor; see mnf_instance_norm_fwd."""
        xc = _dev_f32(x, self.device, "x")
        rc = _dev_f32(residual, self.device, "residual") if residual is not None else None
        if rc is not None and rc.shape != xc.shape:
            raise ValueError(f"residual {tuple(rc.shape)} != x {tuple(xc.shape)}")
        N, Cc, H, W = xc.shape
        y = torch.empty_like(xc)
        _check(self.lib.mnf_instance_norm_fwd(self._h, xc.data_ptr(), _ptr(rc), y.data_ptr(), N * Cc, H * W, mode, eps,
                                              _stream(self.device)), "mnf_instance_norm_fwd")
        return y

    def instance_norm_nhwc(self, x: torch.Tensor, mode: int = 1, residual: Optional[torch.Tensor] = None, eps: float = 1e-5):
        """Fused InstanceNorm2d(+ReLU)(+residual, ReLU) on an fp32 / fp16 [N,C,H,W] tensor in torch.channels_last memory format; see
        mnf_instance_norm_nhwc_fwd.  Returns a channels_last tensor of the same dtype."""
        for t, nm in ((x, "x"), (residual, "residual")):
            if t is None:
                continue
            if t.dtype != x.dtype or x.dtype not in (torch.float16, torch.float32) or t.device != self.device or t.dim() != 4 \
                    or not t.is_contiguous(memory_format=torch.channels_last):
                raise ValueError(f"{nm} must be an fp32 / fp16 channels_last [N,C,H,W] tensor on {self.device}")
        if residual is not None and residual.shape != x.shape:
            raise ValueError(f"residual {tuple(residual.shape)} != x {tuple(x.shape)}")
        N, Cc, H, W = x.shape
        y = torch.empty_like(x, memory_format=torch.channels_last)
        _check(self.lib.mnf_instance_norm_nhwc_fwd(self._h, x.data_ptr(), _ptr(residual), y.data_ptr(), int(x.dtype == torch.float16), N, H * W,
                                                   Cc, mode, eps, _stream(self.device)), "mnf_instance_norm_nhwc_fwd")
        return y

    def token_layernorm(self, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5,
                        residual: Optional[torch.Tensor] = None, prefix: Optional[torch.Tensor] = None) -> torch.Tensor:
        """LayerNorm over the last (128-wide) axis of x (fp32 or fp16) fused with what follows it in TransformerLayer.forward:
        ``residual + LN(x)`` (fp32) or, with ``prefix``, ``cat([prefix, LN(x)], -1)`` in fp16; see mnf_token_layernorm_fwd."""
        if x.dtype not in (torch.float32, torch.float16) or not x.is_cuda or x.device != self.device:
            raise ValueError("x must be an fp32 / fp16 tensor on this context's device")
        xc = x.contiguous()
        if xc.shape[-1] != 128:
            raise ValueError("token_layernorm is built for 128 channels")
        n = xc.numel() // 128
        w, b = _dev_f32(weight, self.device, "weight"), _dev_f32(bias, self.device, "bias")
        rc = _dev_f32(residual, self.device, "residual") if residual is not None else None
        pc = _dev_f32(prefix, self.device, "prefix") if prefix is not None else None
        for t, nm in ((rc, "residual"), (pc, "prefix")):
            if t is not None and t.shape != xc.shape:
                raise ValueError(f"{nm} {tuple(t.shape)} != x {tuple(xc.shape)}")
        if pc is not None and rc is not None:
            raise ValueError("residual and prefix are exclusive")
        if pc is not None:
            out = torch.empty(xc.shape[:-1] + (256,), dtype=torch.float16, device=self.device)
            o32, o16 = None, out.data_ptr()
        else:
            out = torch.empty(xc.shape, dtype=torch.float32, device=self.device)
            o32, o16 = out.data_ptr(), None
        _check(self.lib.mnf_token_layernorm_fwd(self._h, xc.data_ptr(), int(xc.dtype == torch.float16), w.data_ptr(), b.data_ptr(), eps,
                                                _ptr(rc), _ptr(pc), o32, o16, n, 128, _stream(self.device)), "mnf_token_layernorm_fwd")
        return out

    def token_block_pack(self, merge_w, norm1_w, norm1_b, mlp0_w=None, mlp2_w=None, norm2_w=None, norm2_b=None) -> torch.Tensor:
        """Pack one TransformerLayer's post-attention parameters for ``token_block`` (mnf_token_block_pack_weights): fp16 operand
        tiles in streaming order + the LayerNorm vectors.  ``mlp0_w is None`` packs a no_ffn (self-attention) layer."""
        with_ffn = mlp0_w is not None
        ts = [_dev_f32(t.detach(), self.device, nm) if t is not None else None
              for t, nm in ((merge_w, "merge_w"), (norm1_w, "norm1_w"), (norm1_b, "norm1_b"), (mlp0_w, "mlp0_w"), (mlp2_w, "mlp2_w"),
                            (norm2_w, "norm2_w"), (norm2_b, "norm2_b"))]
        if ts[0].shape != (128, 128) or (with_ffn and (ts[3].shape != (1024, 256) or ts[4].shape != (128, 1024))):
            raise ValueError("token_block is built for d_model 128 with ffn_dim_expansion 4")
        blob = torch.empty(self.lib.mnf_token_block_weight_bytes(int(with_ffn)), dtype=torch.uint8, device=self.device)
        _check(self.lib.mnf_token_block_pack_weights(self._h, *[_ptr(t) for t in ts], blob.data_ptr(), _stream(self.device)),
               "mnf_token_block_pack_weights")
        return blob

    def token_block(self, attn_out: torch.Tensor, source: torch.Tensor, blob: torch.Tensor, with_ffn: bool, eps: float = 1e-5) -> torch.Tensor:
        """``source + LN1(merge(attn_out))`` or ``source + LN2(mlp(cat[source, LN1(merge(attn_out))]))`` (models/gmflow/transformer.py:173-185)
        in one tcgen05 kernel; attn_out / source [..., 128] fp32 on this device; see mnf_token_block_fwd."""
        a, s = _dev_f32(attn_out, self.device, "attn_out"), _dev_f32(source, self.device, "source")
        if a.shape != s.shape or a.shape[-1] != 128:
            raise ValueError(f"attn_out {tuple(a.shape)} / source {tuple(s.shape)}: expected equal [..., 128] shapes")
        if blob.numel() != self.lib.mnf_token_block_weight_bytes(int(with_ffn)):
            raise ValueError("packed weights do not match with_ffn")
        out = torch.empty_like(s)
        _check(self.lib.mnf_token_block_fwd(self._h, a.data_ptr(), s.data_ptr(), blob.data_ptr(), int(with_ffn), eps, out.data_ptr(),
                                            a.numel() // 128, 128, _stream(self.device)), "mnf_token_block_fwd")
        return out

    def window_attn(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, h: int, w: int, num_splits: int,
                    with_shift: bool, impl: int = 0, use_workspace: bool = True) -> torch.Tensor:
        q = _dev_f32(q, self.device, "q")
        k = _dev_f32(k, self.device, "k")
        v = _dev_f32(v, self.device, "v")
        B, L, Cc = q.shape
        if L != h * w:
            raise ValueError("q.shape[1] != h*w")
        out = torch.empty_like(q)
        ws = None
        if impl != 1 and use_workspace:
            need = self.lib.mnf_window_attn_workspace_bytes(B, h, w, num_splits)
            ws = torch.empty((max(need, 16),), dtype=torch.uint8, device=self.device)
        _check(self.lib.mnf_window_attn_fwd(self._h, q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, h, w, Cc,
                                            num_splits, int(with_shift), impl, _ptr(ws), ws.numel() if ws is not None else 0,
                                            _stream(self.device)), "mnf_window_attn_fwd")
        return out

    def window_attn_pack_proj(self, wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor) -> torch.Tensor:
        """q_proj / k_proj / v_proj weights ([128,128] fp32, nn.Linear layout) -> packed operand images for ``window_attn_proj``."""
        ws = [_dev_f32(t.detach(), self.device, nm) for t, nm in ((wq, "q_proj"), (wk, "k_proj"), (wv, "v_proj"))]
        if any(t.shape != (128, 128) for t in ws):
            raise ValueError("projection weights must be [128, 128]")
        blob = torch.empty(self.lib.mnf_window_attn_proj_weight_bytes(), dtype=torch.uint8, device=self.device)
        _check(self.lib.mnf_window_attn_pack_proj_weights(self._h, ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr(), blob.data_ptr(),
                                                          _stream(self.device)), "mnf_window_attn_pack_proj_weights")
        return blob

    def window_attn_proj(self, source: torch.Tensor, target: torch.Tensor, blob: torch.Tensor, h: int, w: int, num_splits: int,
                         with_shift: bool, target_roll: int = 0) -> torch.Tensor:
        """``attention(q_proj(source), k_proj(target), v_proj(target))`` (models/gmflow/transformer.py:158-171): projections fused
        into the operand-packing kernel, then the tcgen05 attention kernel; see mnf_window_attn_proj_fwd.  ``target_roll`` r: keys /
        values of batch item b come from ``target[(b + r) % B]``."""
        s_, t_ = _dev_f32(source, self.device, "source"), _dev_f32(target, self.device, "target")
        B, L, Cc = s_.shape
        if L != h * w or t_.shape != s_.shape:
            raise ValueError("source / target must be [B, h*w, 128] with equal shapes")
        out = torch.empty_like(s_)
        need = self.lib.mnf_window_attn_workspace_bytes(B, h, w, num_splits)
        ws = torch.empty((max(need, 16),), dtype=torch.uint8, device=self.device)
        _check(self.lib.mnf_window_attn_proj_fwd(self._h, s_.data_ptr(), t_.data_ptr(), blob.data_ptr(), out.data_ptr(), B, h, w, Cc,
                                                 num_splits, int(with_shift), int(target_roll), ws.data_ptr(), ws.numel(),
                                                 _stream(self.device)),
               "mnf_window_attn_proj_fwd")
        return out

    def image_metrics(self, pred: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None, region=None,
                      data_range: float = 2.0) -> torch.Tensor:
        """PSNR / SSIM sums of a rendered view (mnf_image_metrics_fwd): pred, gt [H,W,3] fp32 on this device, mask [H,W] bool / uint8
        (True = masked out) or None, region = (y0, x0, h, w) or None (whole image).  Returns 4 doubles ON THE DEVICE:
        (squared-error sum, element count, SSIM-map sum, element count)."""
        p, g = _dev_f32(pred, self.device, "pred"), _dev_f32(gt, self.device, "gt")
        if p.dim() != 3 or p.shape[-1] != 3 or g.shape != p.shape:
            raise ValueError(f"pred {tuple(p.shape)} / gt {tuple(g.shape)}: expected equal [H, W, 3] shapes")
        H, W = int(p.shape[0]), int(p.shape[1])
        m = None
        if mask is not None:
            if mask.device != self.device or tuple(mask.shape) != (H, W):
                raise ValueError("mask must be [H, W] on the context's device")
            m = mask.to(torch.uint8).contiguous()
        y0, x0, rh, rw = (0, 0, H, W) if region is None else (int(v) for v in region)
        out = torch.empty(4, dtype=torch.float64, device=self.device)
        _check(self.lib.mnf_image_metrics_fwd(self._h, p.data_ptr(), g.data_ptr(), _ptr(m), H, W, y0, x0, rh, rw, float(data_range),
                                              out.data_ptr(), _stream(self.device)), "mnf_image_metrics_fwd")
        return out

    def selftest_umma(self, a: torch.Tensor, b: torch.Tensor, mode: int) -> torch.Tensor:
        a = a.to(self.device, torch.float16).contiguous()
        b = b.to(self.device, torch.float16).contiguous()
        n = b.shape[1] if mode == 2 else b.shape[0]
        assert a.shape[0] == 128 and a.shape[1] == (b.shape[0] if mode == 2 else b.shape[1])
        d = torch.empty((128, n), dtype=torch.float32, device=self.device)
        _check(self.lib.mnf_selftest_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), n, a.shape[1], mode,
                                          _stream(self.device)), "mnf_selftest_umma")
        return d


_contexts: Dict[int, Context] = {}


def get_context(device=None) -> Context:
    """Process-wide context per device."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _contexts:
        _contexts[idx] = Context(torch.device("cuda", idx))
    return _contexts[idx]
