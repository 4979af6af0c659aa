"""Data-parallel plumbing for the one box of 8 B200s.

The reference is single-process / single-device (pascal_train_darknet.py:57-59); partitioning is new:

  inference  images are independent -> shard the batch across ranks, no collective (`shard_range`).
  training   per-rank batch, per-rank BN statistics, each rank's loss is the mean over its local batch
             (net_utils.py:296), so the global gradient is the MEAN of the rank gradients: one all-reduce(sum)
             per bucket + a 1/world scale.  Gradients live in one flat arena ordered last layer first, so a
             bucket is a contiguous slice that is complete early in the backward pass; `BucketedAllReduce.launch`
             is called right after the kernels producing a bucket are enqueued: NCCL's stream waits for the
             current stream at that point and the collective then overlaps the rest of the backward pass.

Works with any torch.distributed backend (nccl on GPUs; the host logic is tested on CPU with gloo, world size 2).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced [start, end) of `total` items for `rank` (first `total % world` ranks get one more)."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def make_buckets(layer_ranges, bucket_bytes, elem_bytes=4):
    """layer_ranges: [(layer, start, end)] in the order gradients become ready (contiguous, ascending offsets).
    Returns [dict(start, end, ready_after=layer)] with each bucket >= bucket_bytes except possibly the last."""
    buckets, start = [], None
    for idx, (layer, s, e) in enumerate(layer_ranges):
        if start is None:
            start = s
        if (e - start) * elem_bytes >= bucket_bytes or idx == len(layer_ranges) - 1:
            buckets.append(dict(start=start, end=e, ready_after=layer))
            start = None
    return buckets


def bucket_segments(buckets, n_layers):
    """The backward pass cut at the bucket launches: [(first layer, last layer)] in execution (descending layer) order, stretch i
    ending with the layer after which bucket i is complete.  Each stretch is one CUDA graph of the data-parallel step
    (Yolo2Trainer._step_segmented); the all-reduce of bucket i is launched between stretch i and stretch i + 1."""
    segs, hi = [], int(n_layers) - 1
    for b in buckets:
        if not 0 <= b['ready_after'] <= hi:
            raise ValueError('buckets must be ordered by descending ready_after layer')
        segs.append((hi, b['ready_after']))
        hi = b['ready_after'] - 1
    if hi != -1:
        raise ValueError('the last bucket must close at layer 0')
    return segs


class BucketedAllReduce:
    def __init__(self, flat, buckets, group=None, world=None):
        self.flat, self.buckets, self.group = flat, buckets, group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.works = []
        self.next = 0

    def begin(self):
        self.works, self.next = [], 0

    def layer_done(self, layer):
        """Call after the kernels writing `layer`'s gradients are enqueued; launches every bucket that is now complete."""
        if self.world <= 1:
            return
        while self.next < len(self.buckets) and self.buckets[self.next]['ready_after'] == layer:
            b = self.buckets[self.next]
            self.works.append(dist.all_reduce(self.flat[b['start']:b['end']], group=self.group, async_op=True))
            self.next += 1

    def finish(self, scale=True):
        """Wait for all buckets; scale=True turns the sums into means here (one pass over the arena), scale=False leaves the
        SUMS for a consumer that applies 1/world itself (the trainer's update kernel does, y2_adam_step_ex's grad_scale)."""
        if self.world <= 1:
            return
        assert self.next == len(self.buckets), 'not every bucket was launched'
        for w in self.works:
            w.wait()
        if scale:
            self.flat.mul_(1.0 / self.world)
