#!/usr/bin/env python
"""Run the weight-gradient kernels of single Darknet19 layers (batch 64, 416^2 geometry) for ncu / quick timing.
    python tools/run_wgrad.py L1 L2 L19 [--iters 5] [--batch 64]"""
import os
import sys
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from tensorflow_yolo2_b200 import ops  # noqa: E402
from run_layer import layer_specs  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if a.startswith('L')]
    iters = int(sys.argv[sys.argv.index('--iters') + 1]) if '--iters' in sys.argv else 5
    N = int(sys.argv[sys.argv.index('--batch') + 1]) if '--batch' in sys.argv else 64
    specs = layer_specs()
    for name in args:
        k, cin, cout, pool, h, head = specs[name]
        M = N * h * h
        ld_dh = 32 if name == 'L1' else (cout + 63) // 64 * 64
        x = torch.randn((N, h, h, 8 if name == 'L1' else cin), device='cuda').to(torch.bfloat16)
        dh = torch.randn((M, ld_dh), device='cuda').to(torch.bfloat16)
        dw = torch.zeros((k, k, cin, cout), device='cuda')
        run = (lambda: ops.conv_wgrad_c3(x, dh, cout, dw)) if name == 'L1' else (lambda: ops.conv_wgrad_bf16(x, dh, k, cin, cout, dw))
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * k * k * cin * cout * h * h * N
        byts = x.numel() * 2 + dh.numel() * 2
        print('%s wgrad k%d %d->%d %dx%d: %.3f ms  %.1f TFLOP/s  operands %.0f MB -> %.0f GB/s' % (name, k, cin, cout, h, h, ms, fl / ms / 1e9,
                                                                                            byts / 1e6, byts / ms / 1e6), flush=True)


if __name__ == '__main__':
    main()
