"""Multi-GPU plumbing for the renderer: one process per GPU, rays partitioned across ranks, the rendered tiles exchanged with
an IN-PLACE all-gather into persistent buffers the render kernel writes into directly.  ``torch.distributed``
(NCCL on GPUs, gloo in the CPU tests) is the transport; there is no data-path collective besides the exchange step(s):

  * image-parallel (``bench.py`` default, weak scaling): rank r renders target image r of a round of ``world_size`` images, then
    ONE all-gather of the [rays, 5] tiles (``TileGather``);
  * ray-sharded (``MatchNeRF.forward`` with ``shard_over_ranks``, strong scaling / single-image latency, SURVEY 8e route B): the
    encoder's pair batch is split over the ranks (``pair_owner``) and its feature maps all-gathered, then rank r renders a
    contiguous block of image rows (``row_block``) straight into its slot of the ``ImageGather`` buffers, and rgb / depth /
    opacity are all-gathered in place on a side stream.

No buffer is zero-filled, padded or concatenated when the ranks' shares are equal (512 rows over 1/2/4/8 ranks); ragged splits
fall back to one compaction copy.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def row_block(H: int, W: int, rank: int, world: int) -> Tuple[int, int]:
    """(first_ray, n_rays) of the rows rank ``rank`` owns: rows are split as evenly as possible, earlier ranks
    take the remainder, so blocks are contiguous, disjoint and cover the image in row-major order."""
    base, rem = divmod(H, world)
    r0 = rank * base + min(rank, rem)
    rows = base + (1 if rank < rem else 0)
    return r0 * W, rows * W


def pair_owner(n_pairs: int, world: int) -> List[int]:
    """Rank that encodes view pair p (both directions of a pair stay together: the cross-attention of a pair reads its partner's
    tokens in every block, models/gmflow/transformer.py:279-339).  Pairs are dealt round-robin over the first min(world, n_pairs)
    ranks."""
    return [p % min(world, n_pairs) for p in range(n_pairs)]


def gather_tiles(local: torch.Tensor, counts, group=None) -> torch.Tensor:
    """All-gather ragged [n_r, C] tiles (``counts[r]`` rows from rank r) into the full [sum(counts), C] tensor.
    One collective.  Equal counts: the local tile is sent from where it is and the result is the gather buffer itself (no pad, no
    zero fill, no concatenation); ragged counts: tiles are padded to the largest count and compacted afterwards."""
    world = dist.get_world_size(group)
    pad = max(counts)
    if min(counts) == pad:
        out = local.new_empty((world * pad, local.shape[1]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    buf = local.new_empty((pad, local.shape[1]))
    buf[: local.shape[0]] = local
    out = local.new_empty((world * pad, local.shape[1]))
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[r * pad: r * pad + counts[r]] for r in range(world)], dim=0)


class TileGather:
    """Persistent [world, n, C] buffer for the image-parallel round: a rank writes its [n, C] tile into ``local_view()`` and
    ``all_gather()`` exchanges the slots IN PLACE (NCCL: the send buffer is the rank's own slot of the receive buffer)."""

    def __init__(self, n: int, C: int, device, dtype=torch.float32, group=None):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.full = torch.empty((self.world, n, C), dtype=dtype, device=device)

    def local_view(self) -> torch.Tensor:
        return self.full[self.rank]

    def all_gather(self) -> torch.Tensor:
        dist.all_gather_into_tensor(self.full.view(-1), self.full[self.rank].view(-1), group=self.group)
        return self.full


class ImageGather:
    """Persistent rgb [HW,3] / depth [HW] / opacity [HW] buffers of ONE image whose rows are split over the ranks.  The render
    kernel writes a rank's rows straight into ``local_out()`` (views of the buffers); ``all_gather()`` then exchanges the row
    blocks in place -- three back-to-back collectives on a side stream, so the next image's encoder can start underneath.
    ``wait()`` makes the current stream wait for the exchange (call it before reading ``rgb / depth / opacity``)."""

    def __init__(self, H: int, W: int, device, group=None):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.H, self.W = H, W
        self.spans = [row_block(H, W, r, self.world) for r in range(self.world)]
        self.equal = len({n for _, n in self.spans}) == 1
        self.first, self.n = self.spans[self.rank]
        dev = torch.device(device)
        self.rgb = torch.empty((H * W, 3), dtype=torch.float32, device=dev)
        self.depth = torch.empty((H * W,), dtype=torch.float32, device=dev)
        self.opacity = torch.empty((H * W,), dtype=torch.float32, device=dev)
        self.side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.done = None                                         # event of the last exchange (side stream)

    def local_out(self):
        """(rgb [n,3], depth [n], opacity [n]) views of this rank's row block.  The previous exchange must have finished
        reading them: the current stream waits for it here."""
        self.wait()
        a, b = self.first, self.first + self.n
        return self.rgb[a:b], self.depth[a:b], self.opacity[a:b]

    def _exchange(self, full: torch.Tensor, per_ray: int):
        if self.equal:
            a, b = self.first * per_ray, (self.first + self.n) * per_ray
            dist.all_gather_into_tensor(full.view(-1), full.view(-1)[a:b], group=self.group)
        else:                                                    # ragged row split: pad to the largest block, compact afterwards
            pad = max(n for _, n in self.spans) * per_ray
            buf = full.new_empty((pad,))
            buf[: self.n * per_ray] = full.view(-1)[self.first * per_ray: (self.first + self.n) * per_ray]
            out = full.new_empty((self.world * pad,))
            dist.all_gather_into_tensor(out, buf, group=self.group)
            for r, (f, n) in enumerate(self.spans):
                full.view(-1)[f * per_ray: (f + n) * per_ray] = out[r * pad: r * pad + n * per_ray]

    def all_gather(self):
        if self.side is None:                                    # CPU tensors (gloo tests): synchronous
            for t, k in ((self.rgb, 3), (self.depth, 1), (self.opacity, 1)):
                self._exchange(t, k)
            return
        cur = torch.cuda.current_stream(self.rgb.device)
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            for t, k in ((self.rgb, 3), (self.depth, 1), (self.opacity, 1)):
                self._exchange(t, k)
            self.done = torch.cuda.Event()
            self.done.record(self.side)

    def wait(self):
        if self.done is not None:
            torch.cuda.current_stream(self.rgb.device).wait_event(self.done)
            self.done = None


def render_image_sharded(render_range: Callable[..., Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]],
                         H: int, W: int, group=None, buffers: Optional[ImageGather] = None, device=None):
    """Render one H x W target image with its rays sharded over the ranks of ``group``.

    ``render_range(first_ray, n_rays, out)`` renders a contiguous row-major ray range on the calling rank INTO ``out`` =
    (rgb [n,3], depth [n], opacity [n]) -- views of the gather buffers, so nothing is copied before the exchange -- or, if it
    returns tensors instead, those are copied into the views (CPU tests, generic renderers).
    Returns (rgb [HW,3], depth [HW,1], opacity [HW,1]) on every rank: views of ``buffers`` (valid until its next use)."""
    if buffers is None:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and dist.get_backend(group) == "nccl" \
                else torch.device("cpu")
        buffers = ImageGather(H, W, device, group)
    out = buffers.local_out()
    ret = render_range(buffers.first, buffers.n, out)
    if ret is not None:
        for dst, src in zip(out, ret):
            dst.copy_(src.reshape(dst.shape))
    buffers.all_gather()
    buffers.wait()
    return buffers.rgb, buffers.depth[:, None], buffers.opacity[:, None]


# ---------------------------------------------------------------------------------------------------------------- training
class GradBucket:
    """Data-parallel training, one process per GPU (SURVEY 8e "Training"; replaces the reference's nn.DataParallel wrapping,
    coach.py:83-85): every parameter gradient is a VIEW into one persistent flat fp32 buffer, so a step exchanges the gradients
    with ONE ``all_reduce`` (4.77 M parameters = 19 MB; NCCL over NVLink on GPUs, gloo in the CPU tests) and averages them over
    the ranks -- no per-parameter collectives, no flatten / unflatten copies.  ``zero()`` replaces ``optimizer.zero_grad()``
    (the views must survive: ``set_to_none=True`` would drop them)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameter")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("all parameters must share one device and dtype")
        self.group = group
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=dt)
        off = 0
        for p in self.params:
            p.grad = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()
        for p in self.params:                         # an optimizer.zero_grad(set_to_none=True) in between would have dropped the views
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * self.flat.element_size():
                raise RuntimeError("a parameter's .grad no longer lives in the bucket (use GradBucket.zero(), not optimizer.zero_grad())")

    def all_reduce_mean(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(dist.get_world_size(self.group))
        return self.flat


def train_iteration(model, batch, optimizer, bucket: GradBucket, clip_enc: Optional[float] = 1.0, scheduler=None,
                    precision: Optional[str] = None):
    """One training step with the reference's semantics (coach.py:215-243, compute_loss :245-258) across ranks: forward in 'train' mode
    (random rays of this rank's sample), MSE against the target view at those rays, backward into the bucket, ONE gradient all-reduce
    (mean over ranks), ``clip_grad_norm_`` on the encoder AFTER the reduce (coach.py:225-226 clips what the optimiser sees), optimiser
    (and per-iteration scheduler) step.  ``precision``: math mode of the library GEMMs / convolutions of the step, forward and
    backward (``train_path.training_precision``: None = PyTorch's defaults, as the reference runs; "tf32" is faster but opt-in, see
    there).  Returns the local loss."""
    from .train_path import training_precision
    bucket.zero()
    with training_precision(precision):
        pred = model(batch, mode="train")
        images = batch["images"]
        b, _, c = images.shape[:3]
        gt = images[:, -1].reshape(b, c, -1).permute(0, 2, 1)
        if "ray_idx" in pred:
            gt = gt[:, pred["ray_idx"]]
        loss = torch.nn.functional.mse_loss(pred["rgb"], gt)
        loss.backward()
    bucket.all_reduce_mean()
    if clip_enc is not None:
        torch.nn.utils.clip_grad_norm_(model.feat_enc.parameters(), clip_enc)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss.detach()


def steps_per_epoch(n_samples: int, batch_size: int, n_ranks: int) -> int:
    """OneCycleLR length of the reference (coach.py:119: ``len(train_loader) // (batch_size // n_gpus)``) with one process per GPU
    in the place of its ``gpu_ids`` list -- kept as is, quirk included, so learning-rate schedules line up with the reference's."""
    return n_samples // max(batch_size // max(n_ranks, 1), 1)
