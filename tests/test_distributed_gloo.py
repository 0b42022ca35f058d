"""CPU tests (gloo, world_size 2) of the host-side sharding logic of the N > 1 path: contiguous train shards with global
index bases, rank-major all-gather of the packed partials, and the fold that must equal one ascending scan. The per-shard
partials come from the CPU oracle and the fold is restated in numpy here (test infrastructure) — the CUDA partial/merge
kernels themselves are covered by tests/test_gpu_match.py::test_knn2_sharded_merge_equals_full_scan."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fold(parts):
    """numpy restatement of the merge rule; parts: (R, Q) int64 packed partials."""
    p = parts.view(np.uint64)
    B = np.full(p.shape[1], 256, np.int64); S = B.copy(); I = np.full(p.shape[1], 0xffffffff, np.int64)
    for r in range(p.shape[0]):
        b = (p[r] >> np.uint64(48)).astype(np.int64); s = ((p[r] >> np.uint64(32)) & np.uint64(0xffff)).astype(np.int64)
        i = (p[r] & np.uint64(0xffffffff)).astype(np.int64)
        win = (b < B) | ((b == B) & (b < 256) & (i < I))
        S = np.where(win, np.minimum(S, B), np.where(b < S, b, S))
        I = np.where(win, i, I); B = np.where(win, b, B)
        S = np.minimum(S, s)
    return np.where(I == 0xffffffff, -1, I).astype(np.int32), B.astype(np.uint16), S.astype(np.uint16)


def test_shard_range_covers_everything():
    from orb_slam2_refactored_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 10, 10000, 10_000_000):
        for world in (1, 2, 3, 8):
            got = [shard_range(n, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(got[r][1] == got[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in got]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(10_000_000, 8, 3) == (3750000, 5000000)     # configs[4]: 1.25 M train rows per GPU


def _worker(rank, world, port, q, t, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import bindings
    from orb_slam2_refactored_b200.distributed import pack_partial, shard_range
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        o = bindings.Oracle('port')
        b, e = shard_range(len(t), world, rank)
        idx, best, second, _ = o.knn2(q, t[b:e], 50, 0.6)
        gidx = np.where(idx >= 0, idx.astype(np.int64) + b, -1)
        part = torch.from_numpy(pack_partial(best, second, gidx))
        gathered = torch.empty((world, len(q)), dtype=torch.int64)
        dist.all_gather_into_tensor(gathered.view(-1), part)
        if rank == 0:
            np.save(out, gathered.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_scan_equals_full_scan_gloo(tmp_path, oracle_port, world):
    import torch.multiprocessing as mp
    from orb_slam2_refactored_b200 import synth
    q, t = synth.planted_descriptors(21, 400, 5001, dup_every=2)     # duplicates land in different shards: tie order matters
    t[17] = ~q[5]                                                     # a distance-256 row
    out = str(tmp_path / 'gathered.npy')
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, q, t, out), nprocs=world, join=True)
    gathered = np.load(out)
    idx, best, second = fold(gathered)
    widx, wbest, wsecond, _ = oracle_port.knn2(q, t, 50, 0.6)
    assert np.array_equal(idx, widx) and np.array_equal(best, wbest) and np.array_equal(second, wsecond)
