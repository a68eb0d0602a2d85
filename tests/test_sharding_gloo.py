"""CPU, world_size 2, gloo: host logic of the ray partition + tile all-gather (the N>1 path of bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from matchnerf_b200.sharding import gather_tiles, render_image_sharded, row_block


def test_row_blocks_cover_image():
    for H, W, world in ((512, 640, 8), (800, 800, 3), (7, 5, 2), (5, 3, 8)):
        spans = [row_block(H, W, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(n for _, n in spans) == H * W
        for (a, n), (b, _) in zip(spans[:-1], spans[1:]):
            assert a + n == b
        assert all(f % W == 0 and n % W == 0 for f, n in spans)


def _worker(rank, world, port, H, W):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def fake_render(first, n):           # pixel id encoded in the outputs
            ids = torch.arange(first, first + n, dtype=torch.float32)
            return torch.stack([ids, ids * 2, ids * 3], 1), ids + 0.5, ids + 0.25
        rgb, depth, opac = render_image_sharded(fake_render, H, W)
        ids = torch.arange(H * W, dtype=torch.float32)
        assert torch.equal(rgb[:, 1], ids * 2) and torch.equal(depth[:, 0], ids + 0.5) and torch.equal(opac[:, 0], ids + 0.25)
        # ragged gather
        counts = [3, 5][:world]
        t = gather_tiles(torch.full((counts[rank], 2), float(rank)), counts)
        assert t.shape == (sum(counts), 2) and float(t[:3].sum()) == 0 and float(t[3:].mean()) == 1
    finally:
        dist.destroy_process_group()


def test_sharded_render_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 7, 5), nprocs=2, join=True)
