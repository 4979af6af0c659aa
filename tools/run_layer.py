#!/usr/bin/env python
"""Run single Darknet19 layers (batch 64, 416^2 geometry) through the tcgen05 conv, for ncu and for
quick per-layer timing.   python tools/run_layer.py L3 L19 [--iters 5] [--batch 64] [--raw] [--x3]
--raw: float32 conv + bias rows (the batch-statistics path); --x3: the bf16x3 precision mode (hi + lo operand pairs)."""
import os
import sys
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from tensorflow_yolo2_b200 import ops  # noqa: E402
from tensorflow_yolo2_b200.yolo2_nets.darknet import CORE_PLAN  # noqa: E402


def layer_specs(image_size=416, of=125):
    h, out = image_size, {}
    plan = list(CORE_PLAN) + [(3, 1024, 1024, False)] * 3 + [(1, 1024, of, False)]
    for i, (k, cin, cout, pool) in enumerate(plan):
        out['L%d' % (i + 1)] = (k, cin, cout, pool, h, i >= 18)
        if pool:
            h //= 2
    return out


def main():
    args = [a for a in sys.argv[1:] if a.startswith("L")]
    iters = int(sys.argv[sys.argv.index('--iters') + 1]) if '--iters' in sys.argv else 5
    N = int(sys.argv[sys.argv.index('--batch') + 1]) if '--batch' in sys.argv else 64
    specs = layer_specs()
    for name in args:
        k, cin, cout, pool, h, head = specs[name]
        x3 = '--x3' in sys.argv
        cin_p = 2 * cin if x3 else ops.conv_cin_padded(cin)
        x = torch.randn((N, h, h, cin_p), device='cuda').to(torch.bfloat16)
        w = torch.randn((k, k, cin, cout), device='cuda') * 0.05
        wp = ops.pack_weights_bf16_split(w) if x3 else ops.pack_weights_bf16(w)
        scale = torch.ones(cout, device='cuda')
        shift = torch.zeros(cout, device='cuda')
        ld = (cout + 31) // 32 * 32
        kw = dict(scale=scale, shift=shift, leaky=not head, pool=pool, out_f32=head, ldy=ld if head else None)
        if '--raw' in sys.argv:        # float32 conv + bias rows (batch-statistics BN path)
            kw = dict(scale=None, shift=shift, leaky=False, pool=False, out_f32=True, ldy=ld)
        if x3:
            kw.update(split_in=True, split_out=not kw['out_f32'])
        y = ops.conv_fwd_bf16(x, wp, k, cin, cout, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.conv_fwd_bf16(x, wp, k, cin, cout, out=y, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * k * k * cin * cout * h * h * N
        print('%s k%d %d->%d %dx%d pool=%d: %.3f ms  %.1f TFLOP/s' % (name, k, cin, cout, h, h, pool, ms, fl / ms / 1e9),
              flush=True)


if __name__ == '__main__':
    main()
