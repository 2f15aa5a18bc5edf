"""CPU (gloo, world_size 2 and 3): the host-side logic of the SNP-sharded path -- shard ranges, the exactness
of summing per-rank integer Grams, the all-gather of per-rank result slices and the permutation max-reduce."""
import os
import socket

import numpy as np
import pytest

from mixmogam_b200 import parallel


def test_shard_range_covers_and_aligns():
    for m in (1, 127, 128, 129, 1000, 4096, 1000000, 999999):
        for world in (1, 2, 3, 4, 8):
            rs = [parallel.shard_range(m, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == m
            for (b0, e0), (b1, e1) in zip(rs[:-1], rs[1:]):
                assert e0 == b1 and b0 <= e0
            for b, e in rs[:-1]:
                assert (b % 128 == 0 or b == m) and (e % 128 == 0 or e == m)
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) < 256 or m < 128 * world


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, m, n, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.Generator(np.random.PCG64(123))
        snps = rng.binomial(2, 0.3, size=(m, n)).astype(np.int8)
        b, e = parallel.shard_range(m, rank, world)
        x = snps[b:e].astype(np.int64)
        t = np.concatenate([(x >= 1), (x >= 2)], axis=0).astype(np.int32)
        g = torch.from_numpy((t.T @ t).astype(np.int32))
        dist.all_reduce(g, op=dist.ReduceOp.SUM)                      # what allreduce_gram does on the device buffer
        local = np.arange(b, e, dtype=np.float64) * 0.5               # stands for this rank's p-values
        full = parallel.allgather_rows(local, m)
        mx = parallel.allreduce_max(np.array([rank, 10.0 - rank, 3.0]))
        q.put((rank, g.numpy(), full, mx))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_gram_and_gather_gloo(world):
    import torch.multiprocessing as mp
    m, n = 1000, 24
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.Generator(np.random.PCG64(123))
    snps = rng.binomial(2, 0.3, size=(m, n)).astype(np.int64)
    t = np.concatenate([(snps >= 1), (snps >= 2)], axis=0).astype(np.int64)
    ref = t.T @ t
    for rank, g, full, mx in outs:
        assert np.array_equal(g.astype(np.int64), ref)                # integer all-reduce: bit-identical for any world size
        assert np.array_equal(full, np.arange(m) * 0.5)
        assert np.array_equal(mx, np.array([world - 1, 10.0, 3.0]))


def test_split_rows_covers_and_balances():
    from mixmogam_b200 import parallel
    for rows in (1, 7, 9999, 10000):
        for world in (1, 2, 3, 8):
            parts = [parallel.split_rows(rows, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.world_size() == 1


def test_slot_range_covers_the_packed_triangle():
    for slots in (1, 6, 820, 19306):
        for world in (1, 2, 3, 8):
            parts = [parallel.slot_range(slots, r, world) for r in range(world)]
            per = parts[0][2]
            assert all(p[2] == per for p in parts) and per * world >= slots
            assert sum(c for _, c, _ in parts) == slots
            pos = 0
            for b, c, _ in parts:
                assert b == min(pos, b) or c == 0
                if c:
                    assert b == pos
                pos += c
            assert all(b == r * per for r, (b, _, _) in enumerate(parts))
