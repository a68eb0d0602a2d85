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
