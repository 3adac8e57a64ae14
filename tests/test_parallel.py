"""Host logic of the N>1 path on CPU: world_size-2 (and 3) gloo process groups
exercise the camera split, the padded all-gather that carries the per-camera
depth/context features, and the max-over-ranks timing reduction bench.py uses.
(The kernels themselves have no CPU path; the 2-GPU equality check of the
sharded forward is tools/shard_check.py, run under torchrun on the GPU box.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from preworld_b200.parallel import (CameraShard, camera_split, collect_occupancy,
                                    collect_results, reduce_confusion)


def test_camera_split_covers_every_camera_once():
    for n in (1, 5, 6, 7):
        for world in range(1, 10):
            split = camera_split(n, world)
            assert len(split) == world
            cams = [c for s, k in split for c in range(s, s + k)]
            assert cams == list(range(n))
            counts = [k for _, k in split]
            assert max(counts) - min(counts) <= 1
    assert camera_split(6, 2) == [(0, 3), (3, 3)]
    assert camera_split(6, 4) == [(0, 2), (2, 2), (4, 1), (5, 1)]
    assert camera_split(6, 8)[6:] == [(6, 0), (6, 0)]


def test_single_rank_shard_is_identity():
    sh = CameraShard(rank=0, world=1)
    t = torch.arange(24.).view(1, 6, 4)
    assert sh.local_range(6) == (0, 6)
    assert sh.all_gather_cams(t, 6) is t


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_cams, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sh = CameraShard()
        assert (sh.rank, sh.world) == (rank, world)
        c0, cn = sh.local_range(n_cams)
        B, F = 2, 5
        # camera c of sample b carries the value 100*b + c (+ feature index)
        local = torch.empty(B, cn, F)
        for b in range(B):
            for j in range(cn):
                local[b, j] = 100 * b + (c0 + j) + torch.arange(F) / 10
        full = sh.all_gather_cams(local, n_cams)
        want = torch.empty(B, n_cams, F)
        for b in range(B):
            for c in range(n_cams):
                want[b, c] = 100 * b + c + torch.arange(F) / 10
        ok = full.shape == want.shape and torch.equal(full, want)
        # bench.py: time of the job = max over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and t.item() == 10.0 + world - 1
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n_cams', [(2, 6), (3, 7), (4, 6), (8, 6)])
def test_all_gather_cams_gloo(world, n_cams):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cams, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]


def _gather_worker(rank, world, port, size, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        per_rank = -(-size // world)                   # the sampler pads to this
        # sample s sits on rank s % world at position s // world (DistributedSampler)
        ids = [rank + k * world for k in range(per_rank)]
        grids = [torch.full((4, 3, 2), i % 251, dtype=torch.uint8) for i in ids]
        got = collect_occupancy(grids, size)
        dicts = [dict(semantic_occ=[g.numpy()], idx=i) for g, i in zip(grids, ids)]
        got2 = collect_results(dicts, size)
        ok = True
        if rank == 0:
            ok = len(got) == size and all(int(g[0, 0, 0]) == s % 251 for s, g in enumerate(got))
            ok = ok and [d['idx'] for d in got2] == list(range(size))
        else:
            ok = got is None and got2 is None
        hist = torch.full((18, 18), rank + 1, dtype=torch.int64)
        occ = torch.arange(4, dtype=torch.int64) * (rank + 1)
        reduce_confusion([hist, occ])
        tot = world * (world + 1) // 2
        ok = ok and bool((hist == tot).all()) and occ.tolist() == [0, tot, 2 * tot, 3 * tot]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,size', [(2, 7), (3, 9)])
def test_result_gather_gloo(world, size):
    """collect_occupancy / collect_results keep the reference's ordering and truncation
    (mmdet3d/apis/test.py:165-195); reduce_confusion sums the ranks' matrices."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, size, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]


def test_result_gather_single_rank():
    grids = [torch.full((2, 2), i, dtype=torch.uint8) for i in range(3)]
    assert [int(g[0, 0]) for g in collect_occupancy(grids, 2)] == [0, 1]
    assert collect_results([1, 2, 3], 2) == [1, 2]
