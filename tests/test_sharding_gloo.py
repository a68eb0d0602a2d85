"""CPU, world_size 2, gloo: host logic of the ray partition + tile all-gather (the N>1 path of bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from matchnerf_b200.sharding import TileGather, gather_tiles, pair_owner, render_image_sharded, row_block


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
        def fake_render(first, n, out):      # pixel id encoded in the outputs; writes straight into the gather-buffer views
            ids = torch.arange(first, first + n, dtype=torch.float32)
            out[0].copy_(torch.stack([ids, ids * 2, ids * 3], 1))
            out[1].copy_(ids + 0.5)
            out[2].copy_(ids + 0.25)
        rgb, depth, opac = render_image_sharded(fake_render, H, W)
        ids = torch.arange(H * W, dtype=torch.float32)
        assert torch.equal(rgb[:, 1], ids * 2) and torch.equal(depth[:, 0], ids + 0.5) and torch.equal(opac[:, 0], ids + 0.25)
        # ragged gather
        counts = [3, 5][:world]
        t = gather_tiles(torch.full((counts[rank], 2), float(rank)), counts)
        assert t.shape == (sum(counts), 2) and float(t[:3].sum()) == 0 and float(t[3:].mean()) == 1
        # equal shares: in-place exchange, the returned tensor is the gather buffer itself
        e = gather_tiles(torch.full((4, 2), float(rank + 1)), [4] * world)
        assert e.shape == (4 * world, 2) and all(float(e[4 * r: 4 * r + 4].mean()) == r + 1 for r in range(world))
        tg = TileGather(6, 5, "cpu")
        tg.local_view().fill_(float(rank + 7))
        full = tg.all_gather()
        assert all(float(full[r].mean()) == r + 7 for r in range(world))
        assert pair_owner(3, 1) == [0, 0, 0] and pair_owner(3, 2) == [0, 1, 0] and pair_owner(3, 8) == [0, 1, 2]
    finally:
        dist.destroy_process_group()


def test_sharded_render_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 7, 5), nprocs=2, join=True)       # ragged rows (7 over 2 ranks)


def test_sharded_render_gloo_world2_even():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 8, 5), nprocs=2, join=True)       # equal shares: the in-place path
