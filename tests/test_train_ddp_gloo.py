"""CPU, world_size 2, gloo: the data-parallel training step (matchnerf_b200/sharding.py GradBucket / train_iteration): one all-reduce
of a flat gradient buffer, mean over ranks, encoder clipping after the reduce -- two ranks on different samples end with the
parameters a single process gets from the averaged gradient.  A stand-in model with the MatchNeRF interface (feat_enc / nerf_dec,
forward(batch, mode='train') -> rgb at ray_idx) keeps the test on the CPU (the real model has no CPU path)."""
import copy
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from matchnerf_b200.sharding import GradBucket, steps_per_epoch, train_iteration


class TinyNet(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.feat_enc = nn.Sequential(nn.Linear(3, 8), nn.Tanh())
        self.nerf_dec = nn.Linear(8, 3)

    def forward(self, batch, mode=None):
        img = batch["images"][:, 0]                                      # [B, 3, H, W]: "render" the target from source view 0
        b, c = img.shape[:2]
        px = img.reshape(b, c, -1).permute(0, 2, 1)
        idx = batch["pick"]
        batch["ray_idx"] = idx
        batch["rgb"] = self.nerf_dec(self.feat_enc(px[:, idx] * 40.0))   # large activations: the encoder clip must bite
        return batch


def _batch(seed):
    g = torch.Generator().manual_seed(seed)
    return {"images": torch.rand(1, 4, 3, 6, 5, generator=g), "pick": torch.randperm(30, generator=g)[:12]}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = TinyNet()
        bucket = GradBucket(net.parameters())
        opt = torch.optim.AdamW(net.parameters(), lr=1e-2, weight_decay=1e-4)
        for step in range(3):
            train_iteration(net, _batch(10 * step + rank), opt, bucket, clip_enc=0.05)
        flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert torch.equal(gathered[0], gathered[1])                     # ranks stay in lock step
        if rank == 0:
            torch.save(flat, out)
    finally:
        dist.destroy_process_group()


def test_two_ranks_match_single_process_on_averaged_gradients(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "params.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    # single process: average the two ranks' gradients by hand, clip the encoder on the averaged gradient, step
    net = TinyNet()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-2, weight_decay=1e-4)
    for step in range(3):
        grads = []
        for rank in range(2):
            net.zero_grad(set_to_none=True)
            b = net(_batch(10 * step + rank), mode="train")
            gt = b["images"][:, -1].reshape(1, 3, -1).permute(0, 2, 1)[:, b["ray_idx"]]
            nn.functional.mse_loss(b["rgb"], gt).backward()
            grads.append([p.grad.clone() for p in net.parameters()])
        for p, g0, g1 in zip(net.parameters(), *grads):
            p.grad = (g0 + g1) / 2
        torch.nn.utils.clip_grad_norm_(net.feat_enc.parameters(), 0.05)
        opt.step()
    want = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-7), float((got - want).abs().max())


def test_bucket_views_and_guards():
    net = TinyNet()
    bucket = GradBucket(net.parameters())
    assert bucket.flat.numel() == sum(p.numel() for p in net.parameters())
    net(_batch(1), mode="train")["rgb"].sum().backward()
    assert float(bucket.flat.abs().sum()) > 0                            # autograd accumulated INTO the flat buffer
    first = next(net.parameters())
    assert first.grad.data_ptr() == bucket.flat.data_ptr()
    bucket.zero()
    assert float(bucket.flat.abs().sum()) == 0 and float(first.grad.abs().sum()) == 0
    assert bucket.all_reduce_mean() is bucket.flat                       # no process group: a no-op
    torch.optim.SGD(net.parameters(), lr=0.1).zero_grad(set_to_none=True)
    try:
        bucket.zero()
        raise AssertionError("expected the dropped views to be reported")
    except RuntimeError:
        pass
    assert steps_per_epoch(1000, 1, 1) == 1000 and steps_per_epoch(1000, 4, 2) == 500 and steps_per_epoch(1000, 1, 8) == 1000
