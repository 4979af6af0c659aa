#!/usr/bin/env python
"""N-GPU data-parallel equivalence check (run under torchrun): after one training step, every rank's averaged
gradient arena must equal the mean of the gradients each rank's shard produces on its own (computed here by
gathering the un-reduced local gradients), and all ranks must hold identical updated weights."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from tensorflow_yolo2_b200.trainer import Yolo2Trainer  # noqa: E402
from tests.helpers import make_store  # noqa: E402


def main():
    world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    N, IS = 4, 96
    S = IS // 32
    rs = np.random.RandomState(1000 + rank)                      # different shard per rank
    img = torch.tensor(rs.randint(0, 256, (N, IS, IS, 3)).astype(np.uint8))
    lab = np.zeros((N, S, S, 25), dtype=np.float32)
    for n in range(N):
        j, i = rs.randint(0, S, 2)
        lab[n, i, j, 0] = 1
        lab[n, i, j, 1:5] = [(j + 0.5) * 32, (i + 0.5) * 32, 40 + 10 * rank, 50]
        lab[n, i, j, 5 + rs.randint(0, 20)] = 1
    # (a) local, un-reduced gradients: a world-1 trainer on this rank's shard
    st_a, _ = make_store(45, tame=True)
    solo = Yolo2Trainer(N, IS, 45, store=st_a, loss='v1', B=5, device=dev)
    solo.world = 1
    solo.reducer.world = 1
    solo.set_labels(lab)
    solo.in_u8.copy_(img)
    solo.forward(); solo.loss(); solo.backward()
    local_grads = solo.grads.clone()
    # (b) the data-parallel trainer (same initial weights on every rank: same seed)
    st_b, _ = make_store(45, tame=True)
    ddp = Yolo2Trainer(N, IS, 45, store=st_b, loss='v1', B=5, device=dev, bucket_bytes=8 << 20)
    assert ddp.world == world and len(ddp.buckets) > 1
    ddp.set_labels(lab)
    ddp.in_u8.copy_(img)
    ddp.iteration += 1
    ddp._set_lr()
    ddp.forward(); ddp.loss(); ddp.backward()          # the arena now holds the all-reduced SUM (1/world lives in the update kernel)
    torch.cuda.synchronize()
    want = local_grads.clone()
    dist.all_reduce(want)
    err = float((ddp.grads - want).norm() / want.norm())
    ddp.update(_lr_set=True)
    torch.cuda.synchronize()
    assert float(ddp.grads.abs().max()) == 0.0         # cleared behind the read
    # atomics in split-K make bitwise equality impossible; fp32 reduction order noise only
    ok = err < 1e-5
    p = ddp.params.clone()
    pmax = p.clone()
    dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(p, pmax)) or float((p - pmax).abs().max()) < 1e-6
    # (c) the data-parallel step as CUDA graphs (one per stretch between two bucket launches, NCCL eager in between) against
    #     eager launches: three iterations each from the same initial weights
    res = []
    for graph in (False, True):
        st_c, _ = make_store(45, tame=True)
        tr = Yolo2Trainer(N, IS, 45, store=st_c, loss='v1', B=5, device=dev, bucket_bytes=8 << 20, use_cuda_graph=graph)
        tr.set_labels(lab)
        losses = [float(tr.step(img)[4]) for _ in range(3)]
        torch.cuda.synchronize()
        assert (tr.seg_graphs is not None) == graph and tr.iteration == 3
        res.append((losses, tr.params.clone()))
    (l0, p0), (l1, p1) = res
    graph_ok = bool(np.allclose(l0, l1, rtol=2e-2)) and float((p0 - p1).abs().mean()) < 3e-4 and float((p0 - p1).abs().max()) < 1e-2
    pg = p1.clone()
    dist.all_reduce(pg, op=dist.ReduceOp.MAX)
    graph_ok = graph_ok and (bool(torch.equal(p1, pg)) or float((p1 - pg).abs().max()) < 1e-6)
    ok = ok and graph_ok
    flag = torch.tensor([1.0 if (ok and same) else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('ddp_check world=%d buckets=%d grad rel err %.3g weights identical %s, segmented graphs == eager %s (losses %s vs %s) -> %s' %
              (world, len(ddp.buckets), err, same, graph_ok, np.round(l0, 4).tolist(), np.round(l1, 4).tolist(),
               'OK' if flag.item() == 1.0 else 'FAIL'), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == '__main__':
    main()
