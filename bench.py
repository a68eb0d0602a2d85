#!/usr/bin/env python
"""Benchmark of the MatchNeRF per-ray hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--samples S] [--shard images|rays]

A step = one pass of the hot path over one batch of synthetic input: the full forward of ONE 512x640 target view (GMFlow
encoder on the 3 source views incl. K-attn, then K-gather + K-mlp-composite over all 327,680 rays, 64 depth samples --
BASELINE.json configs[1]).  With N > 1 ranks:
  --shard images (default, ``scaling: weak``)  every rank renders its own target view, then ONE all-gather of the tiles;
  --shard rays   (``scaling: strong``)         ONE target view per step for the whole job: the encoder's 6 window-attention
                                               batch items and the rays are split over the ranks (MatchNeRF.forward does this
                                               itself when torch.distributed is initialised), feature maps and rendered tiles
                                               are all-gathered.
``value`` = rays/s over the whole job with inputs resident in HBM; ``e2e`` = the same through ``MatchNeRF.forward`` with the
batch in pinned HOST memory (H2D of the images and D2H of rgb/depth/opacity inside the timed region).

Untimed extras on rank 0 at N = 1: per-kernel CUDA-event timings (``kernels``, ``roofline``), the S = 128 configuration
(``kernels_s128``), ``parity`` of the TIMED workload (encoder + render at 512x640) against the CPU oracle and against the
unmodified reference run in fp32 on the same GPU, ``reference_gpu`` (the unmodified reference's own rays/s on this B200: the
north star's ">= 10x" denominator) and ``cpu_baseline`` (the unmodified reference on the host cores).

``--impl reference`` times the UNMODIFIED reference (baseline/_ref or /root/reference through oracle/reference_shim.py) on
the host cores; the oracle port is used only if the reference tree is absent.
"""
from __future__ import annotations

import argparse
import json
import os

os.environ.setdefault("TQDM_DISABLE", "1")          # the reference wraps its slice loop in tqdm (stderr noise in the logs)
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H_IMG, W_IMG = 512, 640
FLOP_PER_SAMPLE = {64: 262432, 128: 266528}     # decoder, 2*MAC, unpadded
GATHER_BYTES_PER_SAMPLE = 12520                 # 3 views x 4 taps x 512 ch x 2 B + colours + 22 floats out
GATHER_FMA_PER_SAMPLE = 8448                    # 6144 bilinear-blend + 2304 pair-product multiply-adds
ATTN_FLOP_PER_CALL = 4 * 24 * 1280 * 1280 * 128  # 4 L^2 C per window, 24 windows of L = 1280


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d["bf16_tflops"]), tensor_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.25)              # let the first sample land inside the timed region
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def make_opts(S: int, device: str, rand_rays_test: int = 20480):
    from matchnerf_b200.utils import AttrDict
    return AttrDict(dict(
        device=device, n_src_views=3,
        encoder=dict(attn_splits_list=[2], cos_n_group=[2, 8], num_transformer_layers=6, feature_upsampler="network",
                     upsample_factor=2, wo_self_attn=False, feature_sample_local_radius=0, feature_sample_local_dilation=1),
        decoder=dict(net_width=128, net_depth=6, skip=[4], posenc=dict(L_3D=10, L_view=0), raytrans_posenc=False,
                     density_maskfill=False, raytrans_act="ReLU"),
        nerf=dict(legacy_coord=True, wo_render_interval=True, view_dep=True, depth=dict(param="metric"), sample_intvs=S,
                  sample_stratified=True, rand_rays_test=rand_rays_test, rand_rays_train=1024)))


def synthetic_batch(seed: int):
    from oracle import synth
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(1, 4, 3, H_IMG, W_IMG, generator=g)
    extr, intr, nf = synth.synthetic_cameras(H_IMG, W_IMG)
    return dict(images=images, extrinsics=extr, intrinsics=intr, near_fars=nf)


def workload_name(S):
    return f"DTU 3-view 512x640 full-image forward (BASELINE configs[1]), S={S}, random-init weights"


# ================================================================================================ the unmodified reference
def load_reference(S: int, device: str, rand_rays_test: int = 20480):
    """(model, EasyDict, root) of the UNMODIFIED reference with the synthetic weights of this benchmark, or None."""
    from oracle import reference_shim as RS
    from oracle import synth
    ref = RS.reference_root()
    if ref is None:
        return None
    RS.install_shim(ref)
    from models.matchnerf import MatchNeRF as RefNet              # the reference's own class
    opt = RS.reference_options(S, device, ref, **{"nerf.rand_rays_test": rand_rays_test})
    net = RefNet(opt).eval()
    net.feat_enc.load_state_dict(synth.synthetic_encoder(1), strict=True)
    net.nerf_dec.load_state_dict(synth.synthetic_decoder(0), strict=True)
    return net.to(device), RS.EasyDict, ref


def pick_threads(fn):
    """PyTorch's CPU ops do not scale to every core of a large host (128 threads measured 20x slower than 16 on the
    GPU box): time a small probe at a few thread counts and keep the fastest."""
    ncpu = os.cpu_count() or 1
    best, best_t = 1, float("inf")
    for n in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(n)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def reference_cpu_sample(S: int, n_rays: int, steps: int, warmup: int):
    """The reference's own CPU implementation of the path on a bounded sample of the benchmark workload: the encoder on the
    three 512x640 views (timed once) and ``render`` of ``n_rays`` contiguous rays per step.  Returns (rays/s of the whole
    image, ms per step, description dict).  Falls back to the oracle port when the reference tree is absent."""
    batch = synthetic_batch(100)
    hw = H_IMG * W_IMG
    loaded = load_reference(S, "cpu")
    with torch.no_grad():
        if loaded is not None:
            net, ED, ref_root = loaded
            b = ED(images=batch["images"], extrinsics=batch["extrinsics"], intrinsics=batch["intrinsics"], near_fars=batch["near_fars"])
            tgt, ref = net.extract_poses(b)
            imgs = batch["images"][:, :3]
            small = [torch.randn(1, 3, 256, 8, 10), torch.randn(1, 3, 256, 16, 20)]
            pick_threads(lambda: net.render(net.opts, tgt, ray_idx=torch.arange(256), mode="test", ref_poses=ref,
                                            ref_images=imgs[..., :64, :80].contiguous(), ref_feats_list=small))
            net.get_img_feat(imgs[..., :64, :96].contiguous())            # warm-up
            t0 = time.perf_counter()
            feats = net.get_img_feat(imgs)
            t_enc = time.perf_counter() - t0

            def step(i):
                first = (i * 7919 * 640) % (hw - n_rays)
                net.render(net.opts, tgt, ray_idx=torch.arange(first, first + n_rays), mode="test", ref_poses=ref, ref_images=imgs,
                           ref_feats_list=feats)
            kind, what = "reference", f"UNMODIFIED reference ({ref_root}; models/matchnerf.py render + get_img_feat, fp32 torch CPU)"
        else:
            from oracle import encoder_oracle as EO
            from oracle import render_oracle as RO
            from oracle import synth
            enc_sd, dec_sd = synth.synthetic_encoder(1), synth.synthetic_decoder(0)
            imgs3 = batch["images"][0, :3]
            extr, intr, nf = batch["extrinsics"], batch["intrinsics"], batch["near_fars"]
            pf, pi, _ = synth.synthetic_scene(64, 80, seed=1)
            e2, i2, n2 = synth.synthetic_cameras(64, 80)
            pick_threads(lambda: RO.render_rays(dec_sd, RO.to_channels_last(pf), pi[0].permute(0, 2, 3, 1).contiguous(), e2[0, :3, :3],
                                                i2[0, :3], n2[0, :3], e2[0, 3, :3], i2[0, 3], n2[0, 3], torch.arange(256), S))
            EO.encode_views(enc_sd, imgs3[:, :, :64, :96])
            t0 = time.perf_counter()
            feats = EO.encode_views(enc_sd, imgs3)
            t_enc = time.perf_counter() - t0
            fl = RO.to_channels_last([f[None] for f in feats])
            img_l = imgs3.permute(0, 2, 3, 1).contiguous()

            def step(i):
                first = (i * 7919 * 640) % (hw - n_rays)
                RO.render_rays(dec_sd, fl, img_l, extr[0, :3, :3], intr[0, :3], nf[0, :3], extr[0, 3, :3], intr[0, 3], nf[0, 3],
                               torch.arange(first, first + n_rays), S)
            kind, what = "port", "oracle port (the reference tree is absent on this box; fp32 torch CPU)"
        for i in range(warmup):
            step(i)
        t0 = time.perf_counter()
        for i in range(steps):
            step(warmup + i)
        t_s = (time.perf_counter() - t0) / steps
    value = hw / (t_enc + t_s * hw / n_rays)
    sample = (f"{what}: {n_rays} contiguous rays x {S} samples x 3 views per step ({steps} steps after {warmup} warm-up, "
              f"{t_s * 1e3:.0f} ms each), encoder on the three 512x640 views timed once ({t_enc:.2f} s) and amortised over the "
              f"{hw}-ray image; {torch.get_num_threads()} threads (fastest of 8/16/32/64/all on this host)")
    return value, t_s * 1e3, dict(value=value, unit="rays/s", cores=torch.get_num_threads(), kind=kind, sample=sample)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S = args.samples
    value, ms, cb = reference_cpu_sample(S, args.ref_rays, args.steps, args.warmup)
    line = dict(impl="reference", metric="rays/sec (DTU 3-view 512x640, %d depth samples, encoder + render)" % S, value=value,
                unit="rays/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=workload_name(S), views=3, samples=S, image=[H_IMG, W_IMG]),
                cpu_baseline=cb, e2e=dict(value=value, unit="rays/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def reference_on_gpu(dev, S_list, slices):
    """The unmodified reference's forward(mode='test') on this GPU: fp32 eager with PyTorch's stock defaults, the shipped
    slice sizes (configs/test.yaml:8 rand_rays_test 20480; 4096 = the README's low-memory advice).  This is the denominator of
    the north star's ">= 10x the reference's single-GPU PyTorch rays/sec".  Also returns the fp32 image (TF32 off) of the first
    S for the parity check."""
    out, images = {}, {}
    hw = H_IMG * W_IMG
    batch = synthetic_batch(100)
    for S in S_list:
        for rr in slices:
            loaded = load_reference(S, str(dev), rr)
            if loaded is None:
                return None, {}
            net, ED, ref_root = loaded

            def fwd():
                b = ED(images=batch["images"].to(dev), extrinsics=batch["extrinsics"].to(dev), intrinsics=batch["intrinsics"].to(dev),
                       near_fars=batch["near_fars"].to(dev))
                with torch.no_grad():
                    return net(b, mode="test")
            fwd()                                            # warm-up (cuDNN heuristics, allocator)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fwd()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            out[f"S{S}_slice{rr}"] = dict(ms_per_image=ms, rays_per_s=hw / (ms * 1e-3))
            if rr == slices[-1]:
                prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
                torch.backends.cuda.matmul.allow_tf32 = False
                torch.backends.cudnn.allow_tf32 = False
                r = fwd()
                torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
                images[S] = tuple(r[k][0].float().cpu() for k in ("rgb", "depth", "opacity"))
            del net
            torch.cuda.empty_cache()
    out["how"] = (f"unmodified reference ({ref_root}) MatchNeRF.forward(mode='test') on {torch.cuda.get_device_name(dev)}, fp32 eager, "
                  "PyTorch default math modes, same synthetic weights / images / cameras as the timed workload, 1 warm-up + 1 timed image "
                  "per configuration (CUDA events)")
    return out, images


def train_step_times(dev, S=128, n_rays=1024, steps=5, warmup=2):
    """BASELINE configs[4] (configs/train.yaml: one 512x640 triplet per rank, rand_rays_train 1024 random rays, S = 128 stratified,
    MSE, AdamW wd 1e-4, clip_grad_norm 1.0 on the encoder -- coach.py:215-243): ms per training step (forward + backward +
    optimiser) of this repo and of the unmodified reference, both on this GPU from the same weights and batch."""
    from matchnerf_b200.matchnerf import MatchNeRF
    from matchnerf_b200.utils import AttrDict
    from oracle import synth
    host = synthetic_batch(100)

    def loop(net, make_batch, precision=None):
        from matchnerf_b200.train_path import training_precision
        import contextlib
        net.train()
        enc_params = list(net.feat_enc.parameters())
        optim = torch.optim.AdamW(net.parameters(), lr=5e-4, weight_decay=1e-4)
        gt_all = host["images"][0, 3].permute(1, 2, 0).reshape(-1, 3).to(dev)

        def step():
            with (training_precision(precision) if precision else contextlib.nullcontext()):
                out = net(make_batch(), mode="train")
                loss = ((out["rgb"][0] - gt_all[out["ray_idx"]]) ** 2).mean()
                optim.zero_grad()
                loss.backward()
            torch.nn.utils.clip_grad_norm_(enc_params, 1.0)
            optim.step()
            return loss
        for _ in range(warmup):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps, float(loss.detach())

    res = dict(config=f"configs/train.yaml shape: 1 x (3 + 1) x 512x640 images, {n_rays} random rays x S={S} stratified, MSE + AdamW + "
                      f"clip_grad_norm(encoder, 1.0); {steps} steps after {warmup} warm-up, CUDA events")
    opt = make_opts(S, str(dev))
    opt.nerf.rand_rays_train = n_rays
    m = MatchNeRF(opt)
    m.feat_enc.load_state_dict(synth.synthetic_encoder(1))
    m.nerf_dec.load_state_dict(synth.synthetic_decoder(0))
    m.to(dev)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    res["ours_tf32_ms_per_step"], res["ours_tf32_last_loss"] = loop(m, lambda: AttrDict({k: v.to(dev) for k, v in host.items()}), "tf32")
    m.load_state_dict(sd0)
    res["ours_ms_per_step"], res["ours_last_loss"] = loop(m, lambda: AttrDict({k: v.to(dev) for k, v in host.items()}), None)
    res["ours_path"] = ("K-gather forward + backward: this repo's CUDA kernels behind an autograd Function; decoder / ray transformer / "
                        "compositing / encoder: library GEMMs + cuDNN under autograd (matchnerf_b200/train_path.py); ours_ms_per_step: "
                        "PyTorch's default math modes, as the reference runs (the comparison with reference_gpu_ms_per_step); "
                        "ours_tf32_ms_per_step: TF32 tensor-core math for the whole step, forward and backward "
                        "(train_path.training_precision('tf32'), opt-in: see its docstring for the measured gradient agreement)")
    del m
    torch.cuda.empty_cache()
    try:
        loaded = load_reference(S, str(dev))
        if loaded is not None:
            net, ED, _ = loaded
            net.opts.nerf.rand_rays_train = n_rays
            res["reference_gpu_ms_per_step"], res["reference_last_loss"] = loop(net, lambda: ED(**{k: v.to(dev) for k, v in host.items()}))
            res["speedup"] = res["reference_gpu_ms_per_step"] / res["ours_ms_per_step"]
            del net
    except Exception as e:
        res["reference_error"] = f"{type(e).__name__}: {e}"
    torch.cuda.empty_cache()
    return res


# ================================================================================================ parity of the timed workload
def psnr_pair(ours, ref, seed=0):
    """PSNR of both against a common pseudo ground truth = reference + N(0, 0.045^2) (27 dB, the regime where 0.01 dB <=> 2e-3 RMS);
    misc/metrics.py:35-41 formula."""
    ours, ref = ours.double(), ref.double()
    gt = ref + 0.045 * torch.randn(ref.shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32).double()

    def psnr(a):
        return -10.0 * float(torch.log10(((a - gt) ** 2).mean()))
    return psnr(ref), psnr(ours)


def rms(a, b):
    return float(((a.double() - b.double()) ** 2).mean().sqrt())


def parity_of_timed_workload(model_out, S, ref_gpu_images):
    """The workload that is TIMED (encoder + render of the 512x640 target) against (a) the CPU oracle (oracle encoder on the same
    three images, 512 strided rays) and (b) the unmodified reference run in fp32 on the same GPU (all 327,680 rays)."""
    from oracle import encoder_oracle as EO
    from oracle import render_oracle as RO
    from oracle import synth
    rgb, depth, opac = (t.float().cpu() for t in model_out)
    batch = synthetic_batch(100)
    res = dict(case=f"the timed workload: encoder + render of the 512x640 target view, S={S} (bench synthetic_batch(100))", bar_db=0.01)
    with torch.no_grad():
        feats = EO.encode_views(synth.synthetic_encoder(1), batch["images"][0, :3])
        idx = torch.arange(173, H_IMG * W_IMG, 640)[:512]                 # one ray per image row, marching across the columns
        extr, intr, nf = batch["extrinsics"], batch["intrinsics"], batch["near_fars"]
        o = RO.render_rays(synth.synthetic_decoder(0), RO.to_channels_last([f[None] for f in feats]),
                           batch["images"][0, :3].permute(0, 2, 3, 1).contiguous(), extr[0, :3, :3], intr[0, :3], nf[0, :3],
                           extr[0, 3, :3], intr[0, 3], nf[0, 3], idx, S)
    p_ref, p_ours = psnr_pair(rgb[idx], o[0])
    res["vs_cpu_oracle"] = dict(rays=int(idx.numel()), rgb_rms=rms(rgb[idx], o[0]), depth_rms=rms(depth[idx], o[1]),
                                opacity_rms=rms(opac[idx], o[2]), opacity_mean=float(o[2].mean()), psnr_ref_db=p_ref, psnr_ours_db=p_ours,
                                psnr_delta_db=p_ours - p_ref)
    if ref_gpu_images and S in ref_gpu_images:
        r_rgb, r_depth, r_op = ref_gpu_images[S]
        p_ref, p_ours = psnr_pair(rgb, r_rgb)
        res["vs_reference_gpu_fp32"] = dict(rays=int(rgb.shape[0]), rgb_rms=rms(rgb, r_rgb), depth_rms=rms(depth, r_depth),
                                            opacity_rms=rms(opac, r_op), psnr_ref_db=p_ref, psnr_ours_db=p_ours,
                                            psnr_delta_db=p_ours - p_ref)
    worst = max(abs(res[k]["psnr_delta_db"]) for k in ("vs_cpu_oracle", "vs_reference_gpu_fp32") if k in res)
    res["abs_psnr_delta_db_max"] = worst
    res["within_bar"] = bool(worst <= 0.01)
    return res


# ================================================================================================ our arm
def run_ours(args):
    import torch.distributed as dist

    from matchnerf_b200 import capi
    from matchnerf_b200.matchnerf import MatchNeRF
    from matchnerf_b200.sharding import TileGather
    from matchnerf_b200.utils import AttrDict
    from oracle import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shard_rays = args.shard == "rays" and world > 1

    S = args.samples
    hw = H_IMG * W_IMG

    def build_model(S_):
        m = MatchNeRF(make_opts(S_, str(dev))).eval()
        m.feat_enc.load_state_dict(synth.synthetic_encoder(1))
        m.nerf_dec.load_state_dict(synth.synthetic_decoder(0))
        m.to(dev)
        m.shard_over_ranks = shard_rays          # MatchNeRF.forward: split encoder + rays over the ranks of the default group
        if args.chunk:
            m.render_chunk = args.chunk
        return m

    model = build_model(S)
    ctx = capi.get_context(dev)

    # weak: each rank renders its own target view (image-parallel round); strong: every rank holds the SAME batch
    host = synthetic_batch(100 + (0 if shard_rays else rank))
    host = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    out_host = torch.empty((hw, 5), dtype=torch.float32).pin_memory()
    tiles = TileGather(hw, 5, dev) if (world > 1 and not shard_rays) else None

    def render_tile(batch):
        out = model(batch, mode="test")
        if tiles is not None:                                   # weak: the single NCCL all-gather of the rendered tiles
            t = tiles.local_view()
            t[:, :3].copy_(out.rgb[0]); t[:, 3:4].copy_(out.depth[0]); t[:, 4:5].copy_(out.opacity[0])
            tiles.all_gather()
            return t
        return torch.cat([out.rgb[0], out.depth[0], out.opacity[0]], dim=1)

    def step_resident():
        with torch.no_grad():
            return render_tile(AttrDict(resident))

    def step_e2e():
        with torch.no_grad():
            # images go host -> device every step; the three small camera tensors are consumed on the host (they become the
            # by-value mnf_scene struct), so they stay in the host batch -- MatchNeRF.forward accepts them on either side
            b = AttrDict({k: (v.to(dev, non_blocking=True) if k == "images" else v) for k, v in host.items()})
            tile = render_tile(b)
            out_host.copy_(tile, non_blocking=True)
        return tile

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, clocks

    ms_step, clocks = timed(step_resident, args.steps, max(args.warmup, 3), sample_clocks=True)
    ms_e2e, _ = timed(step_e2e, args.steps, 3)
    images_per_step = 1 if shard_rays else world
    value = images_per_step * hw / (ms_step * 1e-3)
    e2e_value = images_per_step * hw / (ms_e2e * 1e-3)
    h2d = host["images"].numel() * host["images"].element_size()
    d2h = out_host.numel() * out_host.element_size()

    def kernel_times(model_, S_):
        """CUDA-event timings of the three kernels groups in isolation (rank 0), on the launching stream."""
        pk = peaks()
        with torch.no_grad():
            b = AttrDict(resident)
            feats = model_._get_img_feat_static(b["images"][:, :3])
            tgt, ref = model_.extract_poses(b)
            scene = model_._packed_scenes(ref, b["images"][:, :3], feats)[0]
            sc = scene.c_scene(tgt["extrinsics"][0], tgt["intrinsics"][0], tgt["near_fars"][0])
            cfg = model_.nerf_dec.decoder_cfg(model_.opts)
            chunk = min(model_.render_chunk, hw)
            c32, c16 = ctx.gather_cossim(sc, S_, first_ray=0, n_rays=chunk, want_f32=False, want_f16=True)

            def ev_time(fn, reps=5):
                fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps

            t_g = ev_time(lambda: ctx.gather_cossim(sc, S_, first_ray=0, n_rays=chunk, want_f32=False, want_f16=True))
            t_d = ev_time(lambda: ctx.decoder_composite(sc, cfg, cond_f16=c16, first_ray=0, n_rays=chunk, impl=2))
            t_e = ev_time(lambda: model_._get_img_feat_static(b["images"][:, :3]), 3)
            g = torch.Generator().manual_seed(0)
            q, k, v = (torch.randn(6, 64 * 80, 128, generator=g).to(dev) for _ in range(3))
            t_a0 = ev_time(lambda: ctx.window_attn(q, k, v, 64, 80, 2, False), 10)
            t_a1 = ev_time(lambda: ctx.window_attn(q, k, v, 64, 80, 2, True), 10)
            # the other two kernels of a transformer layer: fused q/k/v projection + attention, and K-block (merge + LN [+ FFN + LN] + residual)
            mk = lambda *sh: (torch.randn(*sh, generator=g) / sh[-1] ** 0.5).to(dev)
            pblob = ctx.window_attn_pack_proj(mk(128, 128), mk(128, 128), mk(128, 128))
            t_p1 = ev_time(lambda: ctx.window_attn_proj(q, k, pblob, 64, 80, 2, True, target_roll=3), 10)
            ln = lambda: (1 + 0.1 * torch.randn(128, generator=g)).to(dev)
            blob0 = ctx.token_block_pack(mk(128, 128), ln(), ln())
            blob1 = ctx.token_block_pack(mk(128, 128), ln(), ln(), mk(1024, 256), mk(128, 1024), ln(), ln())
            t_b0 = ev_time(lambda: ctx.token_block(q, k, blob0, False), 10)
            t_b1 = ev_time(lambda: ctx.token_block(q, k, blob1, True), 10)
        n_samp = chunk * S_
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        clk = (clocks or {}).get("sm_max_mhz") or 1965.0
        fma_roof_ms = n_samp * GATHER_FMA_PER_SAMPLE / (n_sm * 128 * clk * 1e6) * 1e3
        dec_tfs = n_samp * FLOP_PER_SAMPLE.get(S_, 262432) / (t_d * 1e-3) / 1e12
        att_tfs0, att_tfs1 = (ATTN_FLOP_PER_CALL / (t * 1e-3) / 1e12 for t in (t_a0, t_a1))
        return dict(
            gather_cossim=dict(ms_per_launch=t_g, rays_per_launch=chunk, algorithmic_GBps=n_samp * GATHER_BYTES_PER_SAMPLE / (t_g * 1e-3) / 1e9,
                               bound="CUDA-core FMA / issue (the 39 MB of feature maps are L2-resident: HBM is not the roof)",
                               fma_roof_ms=fma_roof_ms, frac_of_fma_roof=fma_roof_ms / t_g,
                               note="fma roof = 8448 multiply-adds per sample at 128 FMA/clk/SM x SMs x max SM clock; ncu L2/L1 shares in profiles/"),
            decoder_composite=dict(ms_per_launch=t_d, rays_per_launch=chunk, impl="tcgen05", algorithmic_TFLOPs=dec_tfs,
                                   frac_of_tensor_peak=dec_tfs / pk["tensor"], frac_of_tensor_peak_sustained=dec_tfs / pk["tensor_sustained"]),
            window_attn=dict(ms_per_call_noshift=t_a0, ms_per_call_shift=t_a1, algorithmic_TFLOPs_noshift=att_tfs0,
                             algorithmic_TFLOPs_shift=att_tfs1, frac_of_tensor_peak=0.5 * (att_tfs0 + att_tfs1) / pk["tensor"],
                             calls_per_step=12),
            window_attn_fused_projection=dict(ms_per_call_shift=t_p1, projection_and_packing_ms=t_p1 - t_a1 + 0.011,
                                              note="q/k/v projections (tcgen05) inside the operand-packing kernel + K-attn; the unfused call above "
                                                   "includes the 11 us pre-pack pass this replaces together with three cuBLAS GEMMs", calls_per_step=12),
            token_block=dict(ms_per_call_no_ffn=t_b0, ms_per_call_ffn=t_b1, tokens=6 * 64 * 80,
                             algorithmic_TFLOPs_ffn=6 * 64 * 80 * 2 * (128 * 128 + 256 * 1024 + 1024 * 128) / (t_b1 * 1e-3) / 1e12,
                             frac_of_tensor_peak_ffn=6 * 64 * 80 * 2 * (128 * 128 + 256 * 1024 + 1024 * 128) / (t_b1 * 1e-3) / 1e12 / pk["tensor"],
                             note="merge + LayerNorm (+ FFN + LayerNorm) + residual of a TransformerLayer, one tcgen05 kernel; 6 calls of each kind per step"),
            encoder=dict(ms_per_call=t_e, note="CUDA graph: backbone / up-sampler convolutions (cuDNN) + this repo's kernels")), pk

    # ---- untimed extras (rank 0): per-kernel timing, S = 128, parity of the timed workload, reference on this GPU / the host
    roofline, kernels, kernels_s128, parity, ref_gpu, cpu_baseline, train_step = None, {}, None, None, None, None, None
    n_chunks = (hw + min(model.render_chunk, hw) - 1) // min(model.render_chunk, hw)
    launches = args.steps * model.launches_per_image(n_chunks) if hasattr(model, "launches_per_image") else None
    if rank == 0 and not args.quick:
        kernels, pk = kernel_times(model, S)
        kd = kernels["decoder_composite"]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
        if os.path.exists(tp):            # measured once under `ncu --set full`; scaled to this launch size
            tj = json.load(open(tp))
            if "decoder_tc_kernel" in tj:
                traffic = (tj["decoder_tc_kernel"]["dram_read_bytes"] + tj["decoder_tc_kernel"]["dram_write_bytes"]) * kd["rays_per_launch"] / tj["captured_rays"]
        roofline = dict(kernel="decoder_tc_kernel (K-mlp-composite: the kernel the north star sets a roofline target on)", bound="tensor",
                        achieved=kd["algorithmic_TFLOPs"], peak=pk["tensor"], unit="TFLOP/s", frac=kd["frac_of_tensor_peak"], traffic=traffic,
                        peak_source=pk["source"] + ", bf16_tflops (burst: the kernel is timed in isolation)",
                        note="K-gather is reported in `kernels` against the CUDA-core FMA roof: its feature maps are L2-resident, so an HBM "
                             "fraction would be meaningless")
    if rank == 0 and world == 1 and not args.quick:
        with torch.no_grad():
            out = model(AttrDict(resident), mode="test")
            ours_img = (out.rgb[0].clone(), out.depth[0].clone(), out.opacity[0].clone())
        other = 128 if S == 64 else 64
        m2 = build_model(other)
        def step_other():
            with torch.no_grad():
                return m2(AttrDict(resident), mode="test")
        ms2, _ = timed(step_other, 5, 3)
        k2, _ = kernel_times(m2, other)
        with torch.no_grad():
            o2 = m2(AttrDict(resident), mode="test")
            other_img = (o2.rgb[0].clone(), o2.depth[0].clone(), o2.opacity[0].clone())
        kernels_s128 = dict(samples=other, ms_per_step=ms2, value=hw / (ms2 * 1e-3), unit="rays/s", kernels=k2,
                            note=f"same workload at S={other} (configs/base.yaml:48 ships 128; BASELINE's metric is quoted at 64)")
        del m2
        torch.cuda.empty_cache()
        ref_imgs = {}
        if not args.no_reference_gpu:
            try:
                ref_gpu, ref_imgs = reference_on_gpu(dev, (S, other), (4096, 20480))
            except Exception as e:           # never let the reference's problems take the product line down
                ref_gpu = dict(error=f"{type(e).__name__}: {e}")
            if ref_gpu and "error" not in ref_gpu:
                for key, S_, v_ in ((f"S{S}", S, value), (f"S{other}", other, kernels_s128["value"])):
                    best = max(ref_gpu[f"{key}_slice{rr}"]["rays_per_s"] for rr in (4096, 20480))
                    ref_gpu[f"speedup_{key}"] = v_ / best
        parity = parity_of_timed_workload(ours_img, S, ref_imgs)
        parity["other_S"] = parity_of_timed_workload(other_img, other, ref_imgs)
        if not args.no_cpu_baseline:
            _, _, cpu_baseline = reference_cpu_sample(S, 1024, 2, 1)
        if not args.no_train_step:
            try:
                train_step = train_step_times(dev)
            except Exception as e:               # an extra: never take the headline line down
                train_step = dict(error=f"{type(e).__name__}: {e}")

    if rank == 0:
        par = "single GPU" if world == 1 else (
            f"ONE image per step: encoder batch items + rays sharded over {world} ranks, NCCL all-gather of feature maps and tiles"
            if shard_rays else f"image-parallel x{world} + 1 NCCL all-gather of rendered tiles")
        line = dict(metric="rays/sec (DTU 3-view 512x640, %d depth samples, encoder + render)" % S, value=value, unit="rays/s",
                    n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_step, higher_is_better=True,
                    scaling="strong" if shard_rays else "weak", vs_baseline=None,
                    dtype="f16 operands / f32 accumulate (decoder), f16 feature maps, f32 elsewhere", data="synthetic",
                    config=dict(workload=workload_name(S) + (", one target view per step" if shard_rays or world == 1 else ", one target view per rank"),
                                views=3, samples=S, image=[H_IMG, W_IMG], rays_per_step=images_per_step * hw, parallelism=par,
                                l2="per-step working set (conditioning workspace %.1f GB, activations) exceeds the 126 MB L2; no explicit flush"
                                   % (hw * S * 64 / 1e9)),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit="rays/s", ms_per_step=ms_e2e, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=launches, roofline=roofline, kernels=kernels, kernels_s128=kernels_s128, cpu_baseline=cpu_baseline,
                    reference_gpu=ref_gpu, parity=parity, train_step=train_step)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--shard", default="images", choices=["images", "rays"],
                    help="N > 1: 'images' = one target view per rank (weak scaling), 'rays' = one view per step sharded over the ranks (strong)")
    ap.add_argument("--ref-rays", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--no-train-step", action="store_true")
    ap.add_argument("--quick", action="store_true", help="timed region only (no per-kernel / parity / reference extras)")
    ap.add_argument("--chunk", type=int, default=0, help="rays per render launch (default: MatchNeRF.render_chunk)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
