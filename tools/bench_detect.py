#!/usr/bin/env python
"""BASELINE.json configs[2]: region decode + per-class NMS microbench, 13x13x5 anchors x 20 classes (845 boxes per
image), batch 256.  Reports the two-kernel path, the fused kernel and the split pair against the HBM roofline with SURVEY
section 8(d)'s algorithmic bytes (decode 165 620 B + NMS 98 020 B per image).

    python tools/bench_detect.py [--batch 256] [--grid 13] [--shard strong|weak]
    torchrun --nproc-per-node N ... tools/bench_detect.py --shard strong     # the 256 images split over N GPUs, no collective
                                                            --shard weak     # 256 images per GPU
Under torchrun every rank times its shard (device events, L2 flushed) and rank 0 prints the max over ranks."""
import json, os, sys
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from tensorflow_yolo2_b200 import ops  # noqa: E402
from tensorflow_yolo2_b200.yolo2_nets.net_utils import VOC_ANCHORS  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0))))
    shard = sys.argv[sys.argv.index('--shard') + 1] if '--shard' in sys.argv else 'weak'
    NT = int(sys.argv[sys.argv.index('--batch') + 1]) if '--batch' in sys.argv else 256
    if shard == 'strong':
        from tensorflow_yolo2_b200.parallel import shard_range
        a, b = shard_range(NT, rank, world)
        N = b - a
    else:
        N = NT
    S = int(sys.argv[sys.argv.index('--grid') + 1]) if '--grid' in sys.argv else 13
    pk = os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')
    peak = json.load(open(pk))['hbm_gbs'] if os.path.exists(pk) else 6650.0
    g = torch.Generator(device='cpu').manual_seed(1234 + rank)
    net = (2.0 * torch.randn((N, S, S, 125), generator=g)).cuda()
    an = torch.tensor(VOC_ANCHORS).cuda()
    nbox = S * S * 5
    boxes = torch.empty((N, nbox, 4), device='cuda'); scores = torch.empty((N, nbox, 20), device='cuda')
    ki = torch.empty((N, 20, 64), dtype=torch.int32, device='cuda'); kc = torch.empty((N, 20), dtype=torch.int32, device='cuda')
    alg = N * (nbox * 25 * 4 + nbox * 24 * 4) + N * (nbox * 24 * 4 + nbox * 20)       # decode + NMS bytes
    for thr in (0.3, 0.0):
        def two():
            ops.decode_region(net, an, 20, thr, boxes=boxes, scores=scores)
            ops.nms(boxes, scores, thr, 0.45, 64, keep_idx=ki, keep_count=kc)
        def fused():
            ops.detect_fused(net, an, 20, thr, 0.45, 64, boxes=boxes, scores=scores, keep_idx=ki, keep_count=kc)
        def split():
            ops.detect_split(net, an, 20, thr, 0.45, 64, boxes=boxes, scores=scores, keep_idx=ki, keep_count=kc)
        t2, tf, ts = timeit(two), timeit(fused), timeit(split)
        cand = int((scores > 0).sum())
        if world > 1:
            tt = torch.tensor([t2, tf, ts], dtype=torch.float64, device='cuda')
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)               # slowest rank
            t2, tf, ts = [float(v) for v in tt]
            cnt = torch.tensor([cand, N], dtype=torch.int64, device='cuda')
            dist.all_reduce(cnt)
            cand = int(cnt[0])
            alg_all, n_all = alg * int(cnt[1]) // N, int(cnt[1])
        else:
            alg_all, n_all = alg, N
        if rank != 0:
            continue
        best = min(tf, ts)
        print(json.dumps(dict(workload='decode+NMS microbench %dx%dx5x25, batch %d per GPU (%s sharding of %d over %d GPUs), score_thresh %g'
                                       % (S, S, N, shard, NT, world, thr), n_gpus=world, images_total=n_all,
                              images_per_s_total=round(n_all / best * 1e3), aggregate_gbs=round(alg_all / best / 1e6, 1),
                              candidates=cand, kept=int(kc.clamp(min=0).sum()), algorithmic_bytes=alg,
                              two_kernels_us=round(t2 * 1e3, 1), fused_us=round(tf * 1e3, 1), split_us=round(ts * 1e3, 1),
                              split_gbs=round(alg / ts / 1e6, 1), split_frac_of_hbm_peak=round(alg / ts / 1e6 / peak, 3),
                              two_kernels_gbs=round(alg / t2 / 1e6, 1), fused_gbs=round(alg / tf / 1e6, 1),
                              hbm_peak_gbs=peak, fused_frac_of_hbm_peak=round(alg / tf / 1e6 / peak, 3),
                              images_per_s_fused=round(N / tf * 1e3))), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
