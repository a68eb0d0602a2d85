"""Host-side mirror of ``models/matchnerf.py``: same constructor, attributes and method signatures, with the
per-ray work handed to the CUDA library through the C ABI.

    forward(batch, mode, ...)            models/matchnerf.py:32-73
    render(opt, tgt_pose, ray_idx, ...)  models/matchnerf.py:88-143   -> mnf_render_rays_fwd
    render_by_slices(...)                models/matchnerf.py:145-161
    get_img_feat(imgs, ...)              models/matchnerf.py:183-207  -> GMFlow (K-attn inside)
    query_cond_info(points, ...)         models/matchnerf.py:209-293  -> mnf_query_cond_points_fwd (explicit points)
    render_rays(...)                     alias of ``render`` (the name BASELINE.json uses)

PyTorch is the tensor plumbing (device memory, streams, the encoder's conv / linear library calls); there is no
CPU or pure-PyTorch fallback for the per-ray path.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import capi
from .cond_nerf import CondNeRF
from .gmflow import GMFlow
from .utils import AttrDict, get_opt


class MatchNeRF(nn.Module):
    # rays handed to one kernel launch when rendering a full image.  The reference's ``rand_rays_{mode}`` slice
    # size is a memory device (models/matchnerf.py:151-152); rays are independent, so results do not depend on it.
    # One launch per DTU image (327,680 rays, 3.2 GB of conditioning workspace): 22.23 -> 21.95 ms per image vs 81,920.
    render_chunk = 327680

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        self.nerf_setbg_opaque = False
        self.n_src_views = int(get_opt(opts, "n_src_views", 3))
        dev = get_opt(opts, "device", "cuda")
        self.feat_enc = GMFlow(feature_channels=128, num_scales=1, num_head=1, attention_type="swin", ffn_dim_expansion=4,
                               feature_upsampler=get_opt(opts, "encoder.feature_upsampler", "network"),
                               upsample_factor=int(get_opt(opts, "encoder.upsample_factor", 2)),
                               num_transformer_layers=int(get_opt(opts, "encoder.num_transformer_layers", 6)),
                               device=dev).to(dev)
        self.nerf_dec = CondNeRF(opts).to(dev)
        # encoder.feature_sample_local_radius > 0 (models/gmflow/utils.py:136-162): mean over (2r+1)^2 dilated samples, forward only
        self.local_radius = int(get_opt(opts, "encoder.feature_sample_local_radius", 0))
        self.local_dilation = int(get_opt(opts, "encoder.feature_sample_local_dilation", 1))
        if not 0 <= self.local_radius <= 8 or (self.local_radius > 0 and self.local_dilation < 1):
            raise NotImplementedError("feature_sample_local_radius outside 0..8 (dilation >= 1) is not built")
        if not get_opt(opts, "nerf.legacy_coord", True) or get_opt(opts, "nerf.depth.param", "metric") != "metric":
            raise NotImplementedError("only legacy_coord=True with metric depth (all shipped configs) is built")
        self._scene_cache = None

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _unwrap(m):
        return m.module if isinstance(m, nn.DataParallel) else m

    def extract_poses(self, batch):
        """models/matchnerf.py:75-86."""
        tgt = dict(extrinsics=batch["extrinsics"][:, -1, :3, :], intrinsics=batch["intrinsics"][:, -1],
                   near_fars=batch["near_fars"][:, -1])
        ref = dict(extrinsics=batch["extrinsics"][:, :-1, :3, :], intrinsics=batch["intrinsics"][:, :-1],
                   near_fars=batch["near_fars"][:, :-1])
        return tgt, ref

    # Replay the encoder as a CUDA graph in inference (one graph per input shape / device / parameter state): its ~250
    # launches average 20 us of GPU work each, so eager launching leaves the GPU idle between them (5.0 -> 4.5 ms at DTU
    # size).  The returned feature tensors are then the graph's static outputs: valid until the next get_img_feat call.
    encoder_cuda_graph = True

    def get_img_feat(self, imgs, attn_splits_list=None, cur_n_src_views=3) -> List[torch.Tensor]:
        """[B,V,3,H,W] -> [[B,V,256,H/8,W/8], [B,V,256,H/4,W/4]]: view i holds the features it got as a member of
        each of its pairs (models/matchnerf.py:183-207).  Returns FRESH tensors, as the reference does: the CUDA-graph replay
        writes into static buffers that the next call overwrites, so the public method hands out copies (78 MB at DTU size,
        ~25 us); ``forward`` consumes the static buffers directly (they are packed before the next encoder call)."""
        feats = self._get_img_feat_static(imgs, attn_splits_list, cur_n_src_views)
        if getattr(self, "_enc_graph", None) is not None and any(f is g for f in feats for g in self._enc_graph[3]):
            feats = [f.clone() for f in feats]
        return feats

    def _get_img_feat_static(self, imgs, attn_splits_list=None, cur_n_src_views=3) -> List[torch.Tensor]:
        """get_img_feat whose result may alias the encoder graph's static output buffers (valid until the next call)."""
        if attn_splits_list is None:
            attn_splits_list = get_opt(self.opts, "encoder.attn_splits_list", [2])
        world = self._shard_world()
        if world > 1 and imgs.is_cuda and not torch.is_grad_enabled():
            return self._get_img_feat_sharded(imgs, attn_splits_list, cur_n_src_views, world)
        return self._regroup_pairs(self._encode_pairs_static(imgs, attn_splits_list, cur_n_src_views, None), cur_n_src_views)

    def _encode_pairs_static(self, imgs, attn_splits_list, cur_n_src_views, pair_ids):
        """GMFlow on the listed view pairs (None = all): [(f0, f1) per scale], each [B, P', 128, h, w].  CUDA-graph replay in
        inference: the result then lives in the graph's static buffers (valid until the next call with the same key)."""
        if not (self.encoder_cuda_graph and imgs.is_cuda and imgs.dtype == torch.float32 and not torch.is_grad_enabled()
                and not torch.cuda.is_current_stream_capturing()):
            return self._encode_pairs(imgs, attn_splits_list, cur_n_src_views, pair_ids)
        enc = self._unwrap(self.feat_enc)
        params = tuple(enc.parameters())
        from .gmflow import TransformerLayer
        key = (tuple(imgs.shape), imgs.device, tuple(attn_splits_list), cur_n_src_views, enc.matmul_precision,
               str(TransformerLayer.ffn_dtype), bool(TransformerLayer.fused_block), bool(TransformerLayer.fused_proj), bool(getattr(enc, 'token_path', False)), str(getattr(enc.backbone, "fast_dtype", None)), bool(get_opt(self.opts, "encoder.wo_self_attn", False)),
               None if pair_ids is None else tuple(pair_ids), tuple(p.data_ptr() for p in params), sum(p._version for p in params))
        hit = getattr(self, "_enc_graph", None)
        if hit is None or hit[0] != key:
            static_in = imgs.detach().clone()
            side = torch.cuda.Stream(device=imgs.device)
            side.wait_stream(torch.cuda.current_stream(imgs.device))
            with torch.cuda.stream(side):                       # warm-up outside the capture: lazy handles, autotuning, caches
                for _ in range(2):
                    self._encode_pairs(static_in, attn_splits_list, cur_n_src_views, pair_ids)
            torch.cuda.current_stream(imgs.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph):
                    static_out = self._encode_pairs(static_in, attn_splits_list, cur_n_src_views, pair_ids)
            except RuntimeError as e:            # an op that cannot be captured on this torch build: stay eager (same kernels)
                import warnings
                warnings.warn(f"matchnerf_b200: encoder CUDA-graph capture failed ({e}); running the encoder eagerly")
                self.encoder_cuda_graph = False
                torch.cuda.synchronize(imgs.device)
                return self._encode_pairs(imgs, attn_splits_list, cur_n_src_views, pair_ids)
            hit = (key, graph, static_in, static_out)
            self._enc_graph = hit
        hit[2].copy_(imgs)
        hit[1].replay()
        self._feat_epoch = getattr(self, "_feat_epoch", 0) + 1     # static outputs rewritten in place: invalidates packed scenes
        return hit[3]

    def _encode_pairs(self, imgs, attn_splits_list, cur_n_src_views=3, pair_ids=None):
        V = cur_n_src_views
        self._feat_epoch = getattr(self, "_feat_epoch", 0) + 1     # a new set of feature maps: packed scenes of older ones are stale
        out = self.feat_enc(imgs=imgs[:, :V], attn_splits_list=attn_splits_list, keep_raw_feats=True,
                            wo_self_attn=bool(get_opt(self.opts, "encoder.wo_self_attn", False)), pair_ids=pair_ids)
        return list(zip(out["aug_feat0s"], out["aug_feat1s"]))

    @staticmethod
    def _regroup_pairs(pair_feats, V=3) -> List[torch.Tensor]:
        """[(f0, f1) per scale] with all P pairs -> per-view 256-channel maps (models/matchnerf.py:192-205): view i holds the
        features it got as a member of each of its pairs."""
        pairs = [(a, b) for a in range(V - 1) for b in range(a + 1, V)]
        feats = []
        for f0, f1 in pair_feats:
            per_view = [[] for _ in range(V)]
            for p, (i, j) in enumerate(pairs):
                per_view[i].append(f0[:, p])
                per_view[j].append(f1[:, p])
            feats.append(torch.stack([torch.cat(x, dim=1) for x in per_view], dim=1))
        return feats

    def _get_img_feat_eager(self, imgs, attn_splits_list, cur_n_src_views=3) -> List[torch.Tensor]:
        return self._regroup_pairs(self._encode_pairs(imgs, attn_splits_list, cur_n_src_views, None), cur_n_src_views)

    # ------------------------------------------------------------------ multi-GPU: one image split over the ranks (SURVEY 8e route B)
    # Off by default: a model inside an image-parallel job (one target view per rank) must not start collectives of its own.
    # ``MNF_SHARD_RANKS=1`` or ``model.shard_over_ranks = True`` turns it on for torchrun-launched test.py / bench.py --shard rays.
    shard_over_ranks = bool(int(os.environ.get("MNF_SHARD_RANKS", "0")))

    def _shard_world(self) -> int:
        if not self.shard_over_ranks:
            return 1
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _get_img_feat_sharded(self, imgs, attn_splits_list, V, world):
        """Encoder split over the ranks: the CNN backbone runs replicated (3 views), the transformer + up-sampler of view pair p
        only on rank ``pair_owner[p]``; the pair's four feature maps travel in ONE broadcast from a contiguous per-pair buffer."""
        import torch.distributed as dist
        from .sharding import pair_owner
        rank = dist.get_rank()
        n_pairs = V * (V - 1) // 2
        owner = pair_owner(n_pairs, world)
        mine = [p for p in range(n_pairs) if owner[p] == rank]
        B, _, _, H, W = imgs.shape
        local = self._encode_pairs_static(imgs, attn_splits_list, V, mine) if mine else None
        key = (tuple(imgs.shape), imgs.device)
        ex = getattr(self, "_pair_exchange", None)
        if ex is None or ex[0] != key:
            if local is not None:
                shapes = [tuple(f0.shape[2:]) for f0, _ in local]
            else:                                                  # a rank without pairs: shapes from the encoder geometry
                hh, ww = (768, 1024) if (H, W) == (756, 1008) else (H, W)
                up = int(get_opt(self.opts, "encoder.upsample_factor", 2))
                shapes = [(128, hh // 8, ww // 8), (128, hh // 8 * up, ww // 8 * up)]
            sizes = [B * c * h * w for c, h, w in shapes]
            bufs = [torch.empty((2 * sum(sizes),), dtype=torch.float32, device=imgs.device) for _ in range(n_pairs)]
            ex = (key, bufs, shapes, sizes)
            self._pair_exchange = ex
        _, bufs, shapes, sizes = ex

        def views(p):
            out, off = [], 0
            for (c, h, w), n in zip(shapes, sizes):
                out.append((bufs[p][off: off + n].view(B, c, h, w), bufs[p][off + n: off + 2 * n].view(B, c, h, w)))
                off += 2 * n
            return out
        for j, p in enumerate(mine):
            for (d0, d1), (f0, f1) in zip(views(p), local):
                d0.copy_(f0[:, j])
                d1.copy_(f1[:, j])
        for p in range(n_pairs):
            dist.broadcast(bufs[p], src=owner[p])
        self._feat_epoch = getattr(self, "_feat_epoch", 0) + 1
        per_pair = [views(p) for p in range(n_pairs)]
        pair_feats = [(torch.stack([per_pair[p][s][0] for p in range(n_pairs)], 1), torch.stack([per_pair[p][s][1] for p in range(n_pairs)], 1))
                      for s in range(len(shapes))]
        return self._regroup_pairs(pair_feats, V)

    def _render_image_sharded(self, opt, tgt_pose, mode, ref_poses, ref_images, ref_feats_list):
        """One full image with its rows split over the ranks: each rank renders its row block straight into its slot of the
        persistent gather buffers, then rgb / depth / opacity are all-gathered in place (sharding.ImageGather)."""
        from .sharding import ImageGather
        B = ref_images.shape[0]
        H, W = ref_images.shape[-2:]
        key = (H, W, ref_images.device)
        hit = getattr(self, "_image_gather", None)
        if hit is None or hit[0] != key:
            hit = (key, ImageGather(H, W, ref_images.device))
            self._image_gather = hit
        ig = hit[1]
        outs = []
        for b in range(B):
            out = ig.local_out()
            self._render(opt, tgt_pose, None, ig.first, ig.n, mode, ref_poses, ref_images, ref_feats_list, out=out, only_b=b)
            ig.all_gather()
            ig.wait()
            outs.append((ig.rgb.clone(), ig.depth[:, None].clone(), ig.opacity[:, None].clone()))
        return AttrDict(rgb=torch.stack([o[0] for o in outs]), depth=torch.stack([o[1] for o in outs]),
                        opacity=torch.stack([o[2] for o in outs]))

    def launches_per_image(self, n_chunks: int = 1) -> int:
        """Kernels of THIS library launched per full-image forward (bench.py's gpu_launches): tensor-core gather + its v3 fix-up pass
        + decoder per render chunk; per transformer layer (12) the fused projection / operand-packing kernel, K-attn and K-block;
        15 NHWC instance norms of two kernels each; 2 feature-map + 1 image packing kernels."""
        return 3 * n_chunks + 12 * 3 + 15 * 2 + 3

    def _packed_scenes(self, ref_poses, ref_images, ref_feats_list):
        """Pack (once per set of feature maps) the per-batch-item scenes the kernels read."""
        # identity AND version of every tensor the packed scene is derived from (feature maps, images, all three camera
        # tensors), plus the encoder epoch (graph replays rewrite the static feature buffers in place).  The source tensors are
        # kept alive by the cache entry, so an address cannot be recycled by the allocator while its key is live.
        src = (ref_feats_list[0], ref_feats_list[1], ref_images, ref_poses["extrinsics"], ref_poses["intrinsics"], ref_poses["near_fars"])
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in src) + (getattr(self, "_feat_epoch", 0),)
        if self._scene_cache is not None and self._scene_cache[0] == key:
            return self._scene_cache[1]
        ctx = self._unwrap(self.nerf_dec).sync_to_library()
        scenes = []
        for b in range(ref_images.shape[0]):
            scenes.append(ctx.pack_scene([ref_feats_list[0][b], ref_feats_list[1][b]], ref_images[b],
                                         ref_poses["extrinsics"][b], ref_poses["intrinsics"][b], ref_poses["near_fars"][b],
                                         self.local_radius, self.local_dilation))
        self._scene_cache = (key, scenes, src)
        return scenes

    def _host_poses(self, batch):
        """(tgt_pose, ref_poses) as HOST tensors.  Camera tensors that already live on the host cost nothing (callers may
        leave extrinsics / intrinsics / near_fars there: only the images are consumed on the device); device tensors are
        read back once and remembered by identity, so re-rendering from the same resident batch never synchronises."""
        cams = tuple(batch[k] for k in ("extrinsics", "intrinsics", "near_fars"))
        key = tuple((t.data_ptr(), t._version, t.device, tuple(t.shape)) for t in cams)
        hit = getattr(self, "_pose_cache", None)
        if hit is None or hit[0] != key or any(a is not b for a, b in zip(hit[1], cams)):
            tgt, ref = self.extract_poses(batch)
            hit = (key, cams, {k: v.detach().cpu() for k, v in tgt.items()}, {k: v.detach().cpu() for k, v in ref.items()})
            self._pose_cache = hit
        return hit[2], hit[3]

    def _c_scene(self, scene, tgt_pose, b):
        """mnf_scene struct of batch item b for one target camera, built once per (scene, pose) and reused by every slice."""
        key = (id(scene), id(tgt_pose), b) + tuple((tgt_pose[k].data_ptr(), tgt_pose[k]._version) for k in ("extrinsics", "intrinsics", "near_fars"))
        hit = getattr(self, "_c_scene_cache", None)
        if hit is not None and hit[0] == key and hit[1] is tgt_pose:
            return hit[2]
        sc = scene.c_scene(tgt_pose["extrinsics"][b], tgt_pose["intrinsics"][b], tgt_pose["near_fars"][b])
        self._c_scene_cache = (key, tgt_pose, sc)
        return sc

    # ------------------------------------------------------------------ the per-slice pipeline
    def render(self, opt, tgt_pose=None, ray_idx=None, mode=None, ref_poses=None, ref_images=None, ref_feats_list=None):
        """models/matchnerf.py:88-143.  Returns AttrDict(rgb [B,R,3], depth [B,R,1], opacity [B,R,1])."""
        if ray_idx is None:
            H, W = ref_images.shape[-2:]
            return self._render(opt, tgt_pose, None, 0, H * W, mode, ref_poses, ref_images, ref_feats_list)
        return self._render(opt, tgt_pose, ray_idx, 0, ray_idx.numel(), mode, ref_poses, ref_images, ref_feats_list)

    render_rays = render

    def _render(self, opt, tgt_pose, ray_idx, first_ray, n_rays, mode, ref_poses, ref_images, ref_feats_list, out=None, only_b=None):
        """One kernel-side slice: explicit pixel ids (``ray_idx``) or the contiguous range [first_ray, first_ray+n_rays).
        ``out`` = (rgb [R,3], depth [R], opacity [R]) makes the kernels write into caller buffers (batch item ``only_b``)."""
        if tgt_pose is None:
            raise Exception("Must provide tgt_pose.")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training step (coach.py:215-243): K-gather forward / backward kernels behind an autograd Function, the MLP / ray
            # transformer / compositing as library GEMMs under autograd (train_path.py)
            from .train_path import render_rays_train
            if not ref_images.is_cuda:
                raise RuntimeError("matchnerf_b200: the render path only exists as CUDA kernels (no CPU path)")
            idx = ray_idx if ray_idx is not None else torch.arange(first_ray, first_ray + n_rays, device=ref_images.device)
            strat = mode == "train" and bool(get_opt(opt, "nerf.sample_stratified", False))
            res = [render_rays_train(self, opt, tgt_pose, idx, ref_poses, ref_images, ref_feats_list, b, strat)
                   for b in (range(ref_images.shape[0]) if only_b is None else [only_b])]
            return AttrDict(rgb=torch.stack([r[0] for r in res]), depth=torch.stack([r[1] for r in res]),
                            opacity=torch.stack([r[2] for r in res]))
        dec = self._unwrap(self.nerf_dec)
        ctx = dec.sync_to_library()
        cfg = dec.decoder_cfg(opt)
        S = cfg.n_samples
        B = ref_images.shape[0]
        scenes = self._packed_scenes(ref_poses, ref_images, ref_feats_list)
        stratified = mode == "train" and bool(get_opt(opt, "nerf.sample_stratified", False))
        outs = []
        for b in (range(B) if only_b is None else [only_b]):
            sc = self._c_scene(scenes[b], tgt_pose, b)
            jitter = torch.rand(n_rays, S, device=ctx.device) if stratified else None     # matchnerf.py:168-169
            rgb, depth, opac = ctx.render_rays(sc, cfg, ray_idx=ray_idx, first_ray=first_ray, n_rays=n_rays, jitter=jitter,
                                               setbg_opaque=self.nerf_setbg_opaque, out=out)
            outs.append((rgb, depth[:, None], opac[:, None]))
        return AttrDict(rgb=torch.stack([o[0] for o in outs]), depth=torch.stack([o[1] for o in outs]),
                        opacity=torch.stack([o[2] for o in outs]))

    def render_by_slices(self, opt, tgt_pose, mode=None, ref_poses=None, ref_images=None, ref_feats_list=None):
        """models/matchnerf.py:145-161: contiguous row-major ray ranges, concatenated."""
        assert ref_images is not None, "Must provide the reference images for MatchNeRF."
        H, W = ref_images.shape[-2:]
        step = max(int(get_opt(opt, f"nerf.rand_rays_{mode}", 0) or 0), int(self.render_chunk))
        parts = [self._render(opt, tgt_pose, None, c, min(step, H * W - c), mode, ref_poses, ref_images, ref_feats_list)
                 for c in range(0, H * W, step)]
        return AttrDict({k: torch.cat([p[k] for p in parts], dim=1) for k in ("rgb", "depth", "opacity")})

    def query_cond_info(self, point_samples, ref_poses, ref_images, ref_feats_list):
        """models/matchnerf.py:209-293 on explicit world-space points [B,R,S,3] -> dict(feat_info [B,R,S,10],
        color_info [B,R,S,9], mask_info [B,R,S,3]).  (MatchNeRF.render runs the same kernel fused with ray casting.)"""
        if not point_samples.is_cuda:
            raise RuntimeError("matchnerf_b200: the conditioning query only exists as a CUDA kernel (no CPU path)")
        B, R, S, _ = point_samples.shape
        scenes = self._packed_scenes(ref_poses, ref_images, ref_feats_list)
        ctx = self._unwrap(self.nerf_dec).sync_to_library()
        conds = []
        for b in range(B):
            # the target camera is not used by the point query; hand the kernel source view 0 as a placeholder
            sc = scenes[b].c_scene(ref_poses["extrinsics"][b, 0], ref_poses["intrinsics"][b, 0], ref_poses["near_fars"][b, 0])
            conds.append(ctx.query_cond_points(sc, point_samples[b].float())[0].view(R, S, -1))
        cond = torch.stack(conds)
        return {"feat_info": cond[..., :10], "color_info": cond[..., 10:19], "mask_info": cond[..., 19:22]}

    # ------------------------------------------------------------------ entry point used by Coach
    def get_video_rendering_path(self, tgt_pose, ref_poses, mode, n_frames=30, batch=None):
        """models/matchnerf.py:295-325: a list of target-pose dicts, one per video frame.  'interpolate' loops through the
        source cameras; 'spiral' needs ``batch['c2ws_all']`` (all scene cameras) and uses the target near / far."""
        from . import camera_paths
        paths = []
        for b, w2cs in enumerate(ref_poses["extrinsics"]):
            if mode == "interpolate":
                sq = torch.eye(4, dtype=torch.float64).repeat(w2cs.shape[0], 1, 1)
                sq[:, :3, :] = w2cs.detach().to("cpu", torch.float64)
                c2ws = torch.linalg.inv(sq)[:, :3, :].to(torch.float32).numpy()          # float64 inverse, as :302-303
                path = camera_paths.interpolate_path(c2ws, n_frames)
            elif mode == "spiral":
                assert batch is not None, "Must provide all c2ws and near_far for getting spiral rendering path."
                c2ws_all = batch["c2ws_all"][b].detach().cpu().numpy()
                near_far = tgt_pose["near_fars"][b].detach().cpu().numpy().tolist()
                path = camera_paths.spiral_path(c2ws_all, near_far, rads_scale=float(get_opt(self.opts, "nerf.video_rads_scale", 0.1)),
                                                n_frames=n_frames)
            else:
                raise Exception(f"Unknown video rendering path mode {mode}")
            paths.append(torch.linalg.inv(torch.tensor(path))[:, :3].to(torch.float32))      # back to world->camera, :317
        paths = torch.stack(paths, dim=0)                                                     # [B, n_frames, 3, 4]
        return [dict(extrinsics=paths[:, f], intrinsics=tgt_pose["intrinsics"].clone().detach(),
                     near_fars=tgt_pose["near_fars"].clone().detach()) for f in range(paths.shape[1])]

    # ------------------------------------------------------------------ entry point used by Coach
    def forward(self, batch, mode=None, render_video=False, render_path_mode="interpolate"):
        """models/matchnerf.py:32-73: encoder once, then random rays (train), the full image, or -- with ``render_video`` --
        every frame of a camera path (one encoder pass amortised over all frames); appends rgb / depth / opacity (and
        ray_idx in train mode) to ``batch`` and returns it.  Video frames are moved to the host and concatenated on dim 0."""
        V = self.n_src_views
        ref_images = batch["images"][:, :V]
        # The kernels take the cameras as host values inside mnf_scene.  Fetch them BEFORE queueing any GPU work: one
        # device->host read on an idle stream, after which the encoder and every render launch are queued without a
        # host synchronisation in between (a .cpu() per slice made the GPU idle ~1 ms per DTU image).
        tgt_pose, ref_poses = self._host_poses(batch)
        if render_video:
            assert mode in ["test", "val"], f"Do NOT render video in mode {mode}, change to either 'test' or 'val'."
            frames = self.get_video_rendering_path(tgt_pose, ref_poses, render_path_mode,
                                                   int(get_opt(self.opts, "nerf.video_n_frames", 30)), batch)
        else:
            frames = [tgt_pose]
        ref_feats_list = self._get_img_feat_static(ref_images, attn_splits_list=get_opt(self.opts, "encoder.attn_splits_list", [2]),
                                                   cur_n_src_views=V)
        B, _, _, H, W = ref_images.shape
        n_rand = int(get_opt(self.opts, f"nerf.rand_rays_{mode}", 0) or 0)
        collected: Dict[str, list] = {}
        for cur_pose in frames:
            if n_rand and mode in ("train", "test-optim"):
                batch["ray_idx"] = torch.randperm(H * W, device=ref_images.device)[: n_rand // B]
                ret = self.render(self.opts, cur_pose, ray_idx=batch["ray_idx"], mode=mode, ref_poses=ref_poses,
                                  ref_images=ref_images, ref_feats_list=ref_feats_list)
            elif self._shard_world() > 1 and ref_images.is_cuda:
                ret = self._render_image_sharded(self.opts, cur_pose, mode, ref_poses, ref_images, ref_feats_list)
            elif n_rand:
                ret = self.render_by_slices(self.opts, cur_pose, mode=mode, ref_poses=ref_poses, ref_images=ref_images,
                                            ref_feats_list=ref_feats_list)
            else:
                ret = self.render(self.opts, cur_pose, mode=mode, ref_poses=ref_poses, ref_images=ref_images,
                                  ref_feats_list=ref_feats_list)
            for k, v in ret.items():
                collected.setdefault(k, []).append(v.detach().to("cpu", non_blocking=True) if render_video else v)
        if render_video and ref_images.is_cuda:
            torch.cuda.current_stream(ref_images.device).synchronize()      # the per-frame device->host copies
        for k, v in collected.items():
            batch[k] = torch.cat(v, dim=0) if len(v) > 1 else v[0]
        return batch


models_dict = {"matchnerf": MatchNeRF}
