#!/usr/bin/env python
"""Time the fused first-layer kernel alone (batch 64, 416x416), L2 flushed between launches."""
import os, sys, json
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from tensorflow_yolo2_b200 import ops

N, IS = 64, int(os.environ.get('IS', 416))
g = torch.Generator(device='cpu').manual_seed(0)
img = torch.randint(0, 256, (N, IS, IS, 3), dtype=torch.uint8, generator=g).cuda()
w = (torch.randn(3, 3, 3, 32, generator=g) * 0.1).cuda()
scale = torch.ones(32).cuda(); shift = torch.zeros(32).cuda()
wp = ops.pack_weights_conv1_u8(w, scale)
out = torch.empty((N, IS // 2, IS // 2, 32), dtype=torch.bfloat16, device='cuda')
flush = torch.empty((256 << 20,), dtype=torch.uint8, device='cuda')
ts = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv1_u8_pool(img, wp, shift, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = sorted(ts[2:])
t = ts[len(ts) // 2]
byts = N * IS * IS * 3 + out.numel() * 2
print(json.dumps(dict(us=round(t * 1e3, 1), gbs=round(byts / t / 1e6, 1),
                      tflops=round(N * 2 * 27 * 32 * IS * IS / t / 1e9, 1))))
