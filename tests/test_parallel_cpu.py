"""Host-side data-parallel logic on CPU: world-size-2 gloo process group (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tensorflow_yolo2_b200.parallel import BucketedAllReduce, bucket_segments, make_buckets, shard_range


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


def test_bucket_segments_cover_the_backward_pass():
    """The data-parallel step replays one CUDA graph per stretch between two bucket launches: the stretches tile layers
    n-1 .. 0 in order and stretch i ends exactly where bucket i becomes complete."""
    ranges, off = [], 0
    for layer, n in zip(range(9, -1, -1), [1000, 50, 4000, 10, 10, 10, 7000, 64, 64, 128]):
        ranges.append((layer, off, off + n))
        off += n
    b = make_buckets(ranges, bucket_bytes=4 * 3000)
    segs = bucket_segments(b, 10)
    assert len(segs) == len(b) and segs[0][0] == 9 and segs[-1][1] == 0
    assert [s[1] for s in segs] == [x['ready_after'] for x in b]
    for (h0, l0), (h1, l1) in zip(segs, segs[1:]):
        assert h0 >= l0 and h1 == l0 - 1
    assert bucket_segments(make_buckets(ranges, bucket_bytes=1 << 40), 10) == [(9, 0)]          # one bucket: one stretch
    with pytest.raises(ValueError):
        bucket_segments(b[:-1], 10)                                                           # does not reach layer 0
    with pytest.raises(ValueError):
        bucket_segments(list(reversed(b)), 10)


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
        # the graphed data-parallel step calls layer_done only at the END of each stretch: every bucket must still be launched
        flat2 = mine.clone()
        red2 = BucketedAllReduce(flat2, buckets, None, world)
        red2.begin()
        for hi, lo in bucket_segments(buckets, 5):
            red2.layer_done(lo)
        red2.finish()
        ok = ok and torch.allclose(flat2, want, rtol=1e-6, atol=1e-7)
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
