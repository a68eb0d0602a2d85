"""GPU, >= 2 devices (run with `gpurun --gpus 2`): the ray-sharded forward of ONE image over two NCCL ranks -- encoder split by
view pair, rows split over the ranks, in-place all-gather of rgb / depth / opacity (matchnerf_b200/sharding.py, SURVEY 8e route B)
-- must equal the single-GPU forward BIT FOR BIT: the same kernels run on the same data, only on different devices."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, H, W, S, out_path):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from matchnerf_b200.matchnerf import MatchNeRF
        from matchnerf_b200.utils import AttrDict
        from oracle import synth
        from tests.test_host_cpu import make_opts
        opt = make_opts(**{"nerf.sample_intvs": S})
        opt.device = str(dev)
        m = MatchNeRF(opt).eval()
        m.feat_enc.load_state_dict(synth.synthetic_encoder(1))
        m.nerf_dec.load_state_dict(synth.synthetic_decoder(0))
        m.to(dev)
        g = torch.Generator().manual_seed(4)
        images = torch.rand(1, 4, 3, H, W, generator=g)
        extr, intr, nf = synth.synthetic_cameras(H, W)
        batch = lambda: AttrDict(images=images.to(dev), extrinsics=extr.to(dev), intrinsics=intr.to(dev), near_fars=nf.to(dev))
        with torch.no_grad():
            single = m(batch(), mode="test")
            ref = {k: single[k].clone() for k in ("rgb", "depth", "opacity")}
            m.shard_over_ranks = True
            for _ in range(2):                                   # twice: the persistent gather buffers are reused
                sharded = m(batch(), mode="test")
                for k in ref:
                    assert sharded[k].shape == ref[k].shape, (k, sharded[k].shape, ref[k].shape)
                    assert torch.equal(sharded[k], ref[k]), (rank, k, float((sharded[k] - ref[k]).abs().max()))
            m.shard_over_ranks = False
        if rank == 0:
            open(out_path, "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node (gpurun --gpus 2)")
@pytest.mark.parametrize("H,W,S", [(64, 96, 32), (128, 160, 64)])
def test_ray_sharded_forward_equals_single_gpu(tmp_path, H, W, S):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "ok.txt")
    mp.spawn(_worker, args=(2, port, H, W, S, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def _train_worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from matchnerf_b200.matchnerf import MatchNeRF
        from matchnerf_b200.sharding import GradBucket, train_iteration
        from matchnerf_b200.utils import AttrDict
        from oracle import synth
        from tests.test_host_cpu import make_opts
        H, W = 64, 96
        opt = make_opts(**{"nerf.sample_intvs": 16, "nerf.rand_rays_train": 256, "nerf.sample_stratified": True})
        opt.device = str(dev)
        m = MatchNeRF(opt).train()
        m.feat_enc.load_state_dict(synth.synthetic_encoder(1))
        m.nerf_dec.load_state_dict(synth.synthetic_decoder(0))
        m.to(dev)
        bucket = GradBucket(m.parameters())
        optim = torch.optim.AdamW(m.parameters(), lr=5e-4, weight_decay=1e-4)
        extr, intr, nf = synth.synthetic_cameras(H, W)
        losses = []
        for step in range(3):                                    # every rank trains on its own (scene, rays): data parallel
            g = torch.Generator().manual_seed(1000 * step + rank)
            images = torch.rand(1, 4, 3, H, W, generator=g)
            b = AttrDict(images=images.to(dev), extrinsics=extr.to(dev), intrinsics=intr.to(dev), near_fars=nf.to(dev))
            losses.append(float(train_iteration(m, b, optim, bucket, clip_enc=1.0)))
        flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1])                     # one all-reduce per step keeps the replicas identical
        assert all(l == l and l < 1e3 for l in losses)
        if rank == 0:
            open(out_path, "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node (gpurun --gpus 2)")
def test_data_parallel_training_keeps_replicas_identical(tmp_path):
    """SURVEY 8e "Training": one process per GPU, each on its own sample, ONE NCCL all-reduce of the flat gradient bucket per step."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "ok.txt")
    mp.spawn(_train_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
