"""CPU, world_size 2 over gloo: class-axis sharding + all-gather plumbing of os2d_b200.dist (the per-rank head is
replaced by a deterministic stand-in; the kernels themselves are covered by the GPU tests)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from os2d_b200 import dist as bd


def _fake_head_factory(offset):
    class Fake:
        def __init__(self, maps):
            self.ids = torch.tensor([float(m.flatten()[0]) for m in maps])

        def __call__(self, fm):
            B, _, H, W = fm.shape
            n = self.ids.numel()
            base = self.ids.view(1, n, 1, 1, 1) + fm.mean().item() * 0
            loc = base + torch.arange(4.).view(1, 1, 4, 1, 1) + torch.zeros(B, n, 4, H, W)
            score = base * 10 + torch.zeros(B, n, 1, H, W)
            corners = base * 100 + torch.arange(8.).view(1, 1, 8, 1, 1) + torch.zeros(B, n, 8, H, W)
            return loc, score, score, corners
    return Fake


def _worker(rank, world, port, C, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    maps = [torch.full((1, 4, 2, 2), float(i)) for i in range(C)]
    sh = bd.ClassShardedHead(maps, _fake_head_factory(0))
    loc, score, corners = sh(torch.zeros(2, 4, 3, 5))
    ok = True
    for c in range(C):
        ok &= bool((score[:, c] == c * 10).all()) and bool((loc[:, c, 2] == c + 2).all()) and bool((corners[:, c, 7] == c * 100 + 7).all())
    ok &= loc.shape == (2, C, 4, 3, 5) and score.shape == (2, C, 1, 3, 5) and corners.shape == (2, C, 8, 3, 5)
    results[rank] = ok
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_cover_all_classes():
    for C in (1, 5, 100, 1000):
        for world in (1, 2, 3, 8):
            blocks = [bd.shard_bounds(C, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == C
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in blocks) == bd.padded_block(C, world)


def test_class_sharded_head_world2_gloo():
    for C in (5, 4, 1):        # ragged, even, and fewer classes than ranks
        mgr = mp.Manager()
        results = mgr.dict()
        port = _free_port()
        mp.spawn(_worker, args=(2, port, C, results), nprocs=2, join=True)
        assert results[0] and results[1]


def test_single_process_path_without_process_group():
    maps = [torch.full((1, 4, 2, 2), float(i)) for i in range(3)]
    sh = bd.ClassShardedHead(maps, _fake_head_factory(0))
    loc, score, corners = sh(torch.zeros(1, 4, 2, 2))
    assert score[0, 2, 0, 0, 0].item() == 20.0
