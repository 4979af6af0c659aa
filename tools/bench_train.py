#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[4]): Darknet19-YOLO2 416x416, batch 64 per GPU, forward + loss +
backward + (N > 1: bucketed NCCL gradient all-reduce overlapped with backward) + Adam.  One JSON line on rank 0.

    python tools/bench_train.py [--steps K] [--warmup W] [--loss v1|region] [--batch 64] [--image-size 416]
    torchrun --nproc-per-node N ... tools/bench_train.py

Also prints per-phase device times (forward / loss / backward / update) and the weight-gradient kernels' TFLOP/s."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--loss', default='v1')
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--image-size', type=int, default=416)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from tensorflow_yolo2_b200 import ops
    from tensorflow_yolo2_b200.trainer import Yolo2Trainer
    from tests.helpers import make_store

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    N, IS = args.batch, args.image_size
    S = IS // 32
    of = 45 if args.loss == 'v1' else 125
    st, _ = make_store(of, tame=True)
    tr = Yolo2Trainer(N, IS, of, store=st, loss=args.loss, B=5, device=dev)
    rs = np.random.RandomState(rank)
    img = torch.tensor(rs.randint(0, 256, (N, IS, IS, 3)).astype(np.uint8)).to(dev)
    tr.in_u8.copy_(img)
    if args.loss == 'v1':
        lab = np.zeros((N, S, S, 25), dtype=np.float32)
        for n in range(N):
            for _ in range(rs.randint(1, 4)):
                cx, cy = rs.uniform(0, IS, 2)
                w, h = rs.uniform(20, 300, 2)
                j, i = int(cx * S / IS), int(cy * S / IS)
                if lab[n, i, j, 0] == 0:
                    lab[n, i, j, 0] = 1
                    lab[n, i, j, 1:5] = [cx, cy, w, h]
                    lab[n, i, j, 5 + rs.randint(0, 20)] = 1
        tr.set_labels(lab)
    else:
        G = tr.max_gt
        cnt = rs.randint(1, 4, N).astype(np.int32)
        bx = np.zeros((N, G, 4), dtype=np.float32)
        bx[:, :3] = np.stack([rs.uniform(0.05, 0.95, (N, 3)), rs.uniform(0.05, 0.95, (N, 3)), rs.uniform(0.05, 0.7, (N, 3)),
                              rs.uniform(0.05, 0.7, (N, 3))], axis=-1)
        tr.set_ground_truth(bx, rs.randint(0, 20, (N, G)).astype(np.int32), cnt)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        tr.step()
    sync()
    n0 = ops.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    phases = np.zeros(4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ev[0].record(); tr.forward(); ev[1].record(); tr.loss(); ev[2].record(); tr.backward(); ev[3].record(); tr.update(); ev[4].record()
        ev[4].synchronize()
        phases += [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    e1.record()
    sync()
    launches = ops.launch_count() - n0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    if rank == 0:
        flops_img = 3 * 29.9e9 * (IS / 416.0) ** 2
        line = dict(metric='images/sec Darknet19-YOLO2 %dx%d training step (fwd+loss+bwd+allreduce+Adam)' % (IS, IS),
                    value=world * N / (ms * 1e-3), unit='images/s', n_gpus=world, steps=args.steps, ms_per_step=ms,
                    loss=args.loss, dtype='bf16', data='synthetic', scaling='weak',
                    phases_ms=dict(zip(('forward', 'loss', 'backward', 'update'), (phases / args.steps).round(3).tolist())),
                    approx_tflops=N * flops_img / (ms * 1e-3) / 1e12, gpu_launches=int(launches),
                    final_loss=float(tr.terms[4]))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
