"""Host-side data-parallel logic on CPU: world-size-2 gloo process group (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tensorflow_yolo2_b200.parallel import BucketedAllReduce, make_buckets, shard_range


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_covers_batch():
    for total in (0, 1, 7, 64, 257):
        for world in (1, 2, 3, 8):
            got = [shard_range(total, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == total
            for (a0, a1), (b0, b1) in zip(got, got[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1


def test_make_buckets_contiguous_and_ordered():
    ranges, off = [], 0
    for layer, n in zip(range(9, -1, -1), [1000, 50, 4000, 10, 10, 10, 7000, 64, 64, 128]):
        ranges.append((layer, off, off + n))
        off += n
    b = make_buckets(ranges, bucket_bytes=4 * 3000)
    assert b[0]['start'] == 0 and b[-1]['end'] == off and b[-1]['ready_after'] == 0
    for x, y in zip(b, b[1:]):
        assert x['end'] == y['start'] and x['ready_after'] > y['ready_after']
    assert all((x['end'] - x['start']) * 4 >= 4 * 3000 for x in b[:-1])


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n = 10_000
        rs = np.random.RandomState(100 + rank)
        flat = torch.tensor(rs.randn(n).astype(np.float32))
        mine = flat.clone()
        ranges = [(l, s, e) for l, s, e in zip(range(4, -1, -1), range(0, n, 2000), range(2000, n + 1, 2000))]
        buckets = make_buckets(ranges, bucket_bytes=4 * 3500)
        red = BucketedAllReduce(flat, buckets, None, world)
        red.begin()
        for layer in range(4, -1, -1):        # backward order
            red.layer_done(layer)
        red.finish()
        gathered = [torch.zeros(n) for _ in range(world)]
        dist.all_gather(gathered, mine)
        want = sum(gathered) / world
        ok = torch.allclose(flat, want, rtol=1e-6, atol=1e-7)
        # batch sharding: the ranks' shards tile the global batch
        lo, hi = shard_range(13, rank, world)
        idx = torch.zeros(13)
        idx[lo:hi] = 1
        dist.all_reduce(idx)
        ok = ok and bool((idx == 1).all())
        q.put((rank, ok, len(buckets)))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] >= 2            # more than one bucket exercised
