#!/usr/bin/env python
"""HBM-bound kernels of the training path at Darknet19 layer shapes (batch 64, 416^2): GB/s vs the measured copy peak."""
import json, os, sys
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from tensorflow_yolo2_b200 import ops  # noqa: E402

SHAPES = [('L1', 416, 32, True), ('L2', 208, 64, True), ('L3', 104, 128, False), ('L9', 26, 512, False), ('L19', 13, 1024, False)]


def timeit(fn, iters=10):
    fn(); torch.cuda.synchronize()
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device='cuda')
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    N = 64
    peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) else 6650.0
    for name, H, C, pool in SHAPES:
        M = N * H * H
        ld = (C + 31) // 32 * 32
        ld_dh = max(32, (C + 63) // 64 * 64) if name != 'L1' else 32
        h = torch.randn((M, ld), device='cuda')
        gamma, beta = torch.rand(C, device='cuda') + 0.5, torch.randn(C, device='cuda')
        Ho = H // 2 if pool else H
        dy = torch.randn((N, Ho, Ho, C), device='cuda').to(torch.bfloat16)
        mean, var = ops.bn_stats(h, C, ld=ld)
        scale, shift = ops.bn_fold(gamma, beta, torch.zeros(C, device='cuda'), var, None)
        out = torch.empty((N, Ho, Ho, C), dtype=torch.bfloat16, device='cuda')
        dh = torch.empty((M, ld_dh), dtype=torch.bfloat16, device='cuda')
        dg, db = torch.empty(C, device='cuda'), torch.empty(C, device='cuda')
        t1 = timeit(lambda: ops.bn_stats(h, C, ld=ld, mean=mean, var=var))
        t2 = timeit(lambda: ops.affine_leaky_pool(h, N, H, H, C, ldx=ld, sub=mean, scale=scale, shift=shift, leaky=True, pool=pool, out_bf16=True, out=out))
        t3 = timeit(lambda: ops.bn_leaky_pool_bwd(h, dy, mean, var, gamma, beta, N, H, H, C, ldh=ld, pool=pool, ld_dh=ld_dh, dgamma=dg, dbeta=db, dh=dh))
        b1 = M * C * 4
        b2 = M * C * 4 + out.numel() * 2
        b3 = 2 * M * C * 4 + 2 * dy.numel() * 2 + M * ld_dh * 2
        print('%-4s M=%8d C=%4d pool=%d | bn_stats %.3f ms %.0f GB/s (%.2f) | affine %.3f ms %.0f GB/s (%.2f) | bn_bwd %.3f ms %.0f GB/s (%.2f)' %
              (name, M, C, pool, t1, b1 / t1 / 1e6, b1 / t1 / 1e6 / peak, t2, b2 / t2 / 1e6, b2 / t2 / 1e6 / peak, t3, b3 / t3 / 1e6, b3 / t3 / 1e6 / peak), flush=True)


if __name__ == '__main__':
    main()
