"""Per-layer device time and achieved HBM bandwidth of the memory-bound kernels of the training step (batch statistics, affine +
leaky (+ pool), BN/leaky/pool backward) at the training shapes of BASELINE configs[4] -- the evidence behind DESIGN section 7's
"elementwise" rows.  python tools/bench_train_elementwise.py [--batch 24] [--image-size 416]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from tensorflow_yolo2_b200 import ops                                   # noqa: E402
from tensorflow_yolo2_b200.engine import create_variables              # noqa: E402
from tensorflow_yolo2_b200.variables import VariableStore               # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device='cuda')
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0.record(); fn(); e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=24)
    ap.add_argument('--image-size', type=int, default=416)
    ap.add_argument('--layers', type=str, default='')
    a = ap.parse_args()
    N, IS = a.batch, a.image_size
    layers = create_variables(VariableStore(seed=0), 30)
    f32 = dict(dtype=torch.float32, device='cuda')
    H = IS
    tot = dict(stats=0.0, affine=0.0, bwd=0.0)
    totb = dict(stats=0.0, affine=0.0, bwd=0.0)
    only = [int(x) for x in a.layers.split(',')] if a.layers else None
    print('layer      M     C  | stats us  GB/s | affine us  GB/s | bn bwd us  GB/s')
    for li, L in enumerate(layers):
        C, pool = L['cout'], L['pool']
        Ho = H // 2 if pool else H
        if only is not None and (li + 1) not in only:
            H = Ho
            continue
        last = li == len(layers) - 1
        M = N * H * H
        ldh = (C + 31) // 32 * 32
        ld_dh = 32 if li == 0 else (C + 63) // 64 * 64
        raw = torch.randn((M, ldh), **f32)
        gamma, beta = torch.rand((C,), **f32) + 0.5, torch.randn((C,), **f32)
        mean, var, scale, shift = (torch.empty((C,), **f32) for _ in range(4))
        mm, mv = torch.zeros((C,), **f32), torch.ones((C,), **f32)
        ws = torch.empty((max(ops.bn_stats_workspace_bytes(M, C), ops.bn_bwd_workspace_bytes(M, C)),), dtype=torch.uint8, device='cuda')
        act = torch.empty((N, Ho, Ho, C), dtype=torch.float32 if last else torch.bfloat16, device='cuda')
        dy = torch.randn((N, Ho, Ho, C), **f32).to(torch.float32 if last else torch.bfloat16)
        dh = torch.empty((M, ld_dh), dtype=torch.bfloat16, device='cuda')
        dg, db = torch.empty((C,), **f32), torch.empty((C,), **f32)
        t_s = timed(lambda: ops.bn_stats_fold_train(raw, C, gamma, beta, mm, mv, ld=ldh, workspace=ws, mean=mean, var=var,
                                                    scale=scale, shift=shift))
        t_a = timed(lambda: ops.affine_leaky_pool(raw, N, H, H, C, ldx=ldh, sub=mean, scale=scale, shift=shift, leaky=True,
                                                  pool=pool, out_bf16=not last, out=act))
        t_b = timed(lambda: ops.bn_leaky_pool_bwd(raw, dy, mean, var, gamma, beta, N, H, H, C, ldh=ldh, leaky=True, pool=pool,
                                                  ld_dh=ld_dh, dgamma=dg, dbeta=db, dh=dh, workspace=ws))
        b_raw = M * C * 4
        b_s = b_raw
        b_a = b_raw + act.numel() * act.element_size()
        b_b = 2 * b_raw + 2 * dy.numel() * dy.element_size() + M * C * 2          # reduce pass + apply pass
        print('L%-2d %9d %5d | %8.1f %5.0f | %9.1f %5.0f | %9.1f %5.0f' % (li + 1, M, C, t_s, b_s / t_s / 1e3, t_a, b_a / t_a / 1e3,
                                                                       t_b, b_b / t_b / 1e3))
        for k, t, b in (('stats', t_s, b_s), ('affine', t_a, b_a), ('bwd', t_b, b_b)):
            tot[k] += t
            totb[k] += b
        del raw, act, dy, dh
        H = Ho
    for k in tot:
        print('%-7s total %8.1f us  %6.1f MB  %5.0f GB/s' % (k, tot[k], totb[k] / 1e6, totb[k] / max(tot[k], 1e-9) / 1e3))


if __name__ == '__main__':
    main()
