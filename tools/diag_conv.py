#!/usr/bin/env python
"""GPU diagnostic for the tcgen05 convolution: runs a list of shape cases against a torch fp32
reference on bf16-rounded operands and reports error patterns.  Used during bring-up under gpurun;
the judged parity tests live in tests/."""
import os
import sys
import json
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from tensorflow_yolo2_b200 import ops  # noqa: E402

CASES = {
    # name: (N, H, W, Cin, Cout, k, pool, out_f32, force_tiled)
    'i1x1_64_128': (1, 16, 16, 64, 128, 1, False, False, False),
    't1x1_64_128': (1, 16, 16, 64, 128, 1, False, False, True),
    'i3x3_64_128': (1, 16, 16, 64, 128, 3, False, False, False),
    't3x3_64_128': (1, 16, 16, 64, 128, 3, False, False, True),
    'i3x3_64_128_odd': (3, 13, 13, 64, 128, 3, False, False, False),
    'i3x3_128_64': (2, 26, 26, 128, 64, 3, False, False, False),
    'i3x3_32_64': (2, 16, 16, 32, 64, 3, False, False, False),
    't3x3_32_64_pool': (2, 16, 16, 32, 64, 3, True, False, False),
    'i3x3_3_32': (2, 16, 16, 3, 32, 3, False, False, False),
    't3x3_3_32_pool': (2, 32, 32, 3, 32, 3, True, False, False),
    't3x3_64_128_pool': (2, 24, 24, 64, 128, 3, True, False, False),
    't3x3_256_512_pool26': (4, 26, 26, 256, 512, 3, True, False, False),
    'i3x3_512_1024': (4, 13, 13, 512, 1024, 3, False, False, False),
    'i1x1_1024_512': (4, 13, 13, 1024, 512, 1, False, False, False),
    'i1x1_1024_125_f32': (4, 13, 13, 1024, 125, 1, False, True, False),
    'i1x1_1024_30_f32': (4, 7, 7, 1024, 30, 1, False, True, False),
    'i3x3_1024_1024_f32': (8, 13, 13, 1024, 1024, 3, False, True, False),
    'p3x3_64_128': (2, 64, 72, 64, 128, 3, False, False, False),
    'p3x3_64_128_pool': (2, 64, 64, 64, 128, 3, True, False, False),
    'p3x3_128_64': (2, 80, 64, 128, 64, 3, False, False, False),
    'p3x3_32_64_pool': (2, 96, 64, 32, 64, 3, True, False, False),
    'p3x3_3_32_pool': (2, 64, 96, 3, 32, 3, True, False, False),
    'p3x3_3_32': (1, 70, 66, 3, 32, 3, False, False, False),
    'p3x3_256_512': (1, 64, 64, 256, 512, 3, False, False, False),
    'big_l2': (8, 208, 208, 32, 64, 3, True, False, False),
    'big_l1': (8, 416, 416, 3, 32, 3, True, False, False),
    'big_l3': (8, 104, 104, 64, 128, 3, False, False, False),
}


def run_case(name, spec, seed=0):
    N, H, W, Cin, Cout, k, pool, out_f32, force_tiled = spec
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn((N, H, W, Cin), device='cuda', generator=g)
    w = torch.randn((k, k, Cin, Cout), device='cuda', generator=g) * (1.0 / (k * k * Cin) ** 0.5)
    scale = torch.rand((Cout,), device='cuda', generator=g) + 0.5
    scale[::3] *= -1.0                      # negative gammas: pool must come after the affine
    shift = torch.randn((Cout,), device='cuda', generator=g) * 0.1
    cin_p = ops.conv_cin_padded(Cin)
    xb = torch.zeros((N, H, W, cin_p), device='cuda', dtype=torch.bfloat16)
    xb[..., :Cin] = x.to(torch.bfloat16)
    wp = ops.pack_weights_bf16(w.contiguous())
    if force_tiled:
        os.environ['Y2_CONV_FORCE_TILED'] = '1'
    else:
        os.environ.pop('Y2_CONV_FORCE_TILED', None)
    ldy = ((Cout + 31) // 32 * 32) if out_f32 else None
    torch.cuda.synchronize()
    t0 = time.time()
    y = ops.conv_fwd_bf16(xb, wp, k, Cin, Cout, scale=scale, shift=shift, leaky=not out_f32, pool=pool,
                          out_f32=out_f32, ldy=ldy)
    torch.cuda.synchronize()
    dt = time.time() - t0
    # reference: fp32 conv on the bf16-rounded operands
    xr = x.to(torch.bfloat16).float().permute(0, 3, 1, 2)
    wr = w.to(torch.bfloat16).float().permute(3, 2, 0, 1)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = F.conv2d(xr, wr, padding=k // 2).permute(0, 2, 3, 1)
    ref = ref * scale + shift
    if not out_f32:
        ref = torch.maximum(ref, 0.1 * ref)
    if pool:
        ref = ref.reshape(N, H // 2, 2, W // 2, 2, Cout).amax(dim=(2, 4))
    if out_f32:
        got = y.reshape(N, H, W, -1)[..., :Cout]
    else:
        got = y.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    maxerr = err.max().item()
    tol = 2e-2 * denom if not out_f32 else 2e-3 * denom
    ok = bool(torch.isfinite(got).all().item()) and maxerr <= tol
    info = dict(case=name, ok=ok, maxerr=maxerr, refmax=denom, ms=dt * 1e3)
    if not ok:
        bad = err > tol
        rows_bad = bad.reshape(-1, Cout).any(dim=1)
        cols_bad = bad.reshape(-1, Cout).any(dim=0)
        info['frac_bad'] = bad.float().mean().item()
        info['rows_bad_first'] = torch.nonzero(rows_bad).flatten()[:24].tolist()
        info['n_rows_bad'] = int(rows_bad.sum().item())
        info['n_rows'] = int(rows_bad.numel())
        info['cols_bad_first'] = torch.nonzero(cols_bad).flatten()[:24].tolist()
        info['n_cols_bad'] = int(cols_bad.sum().item())
        info['sample_got'] = got.reshape(-1, Cout)[:4, :6].tolist()
        info['sample_ref'] = ref.reshape(-1, Cout)[:4, :6].tolist()
        info['nan'] = int((~torch.isfinite(got)).sum().item())
    return info


def main():
    names = sys.argv[1:] or list(CASES.keys())
    out = []
    for n in names:
        try:
            info = run_case(n, CASES[n])
        except Exception as e:  # noqa: BLE001
            info = dict(case=n, ok=False, exception=repr(e)[:500])
            print(json.dumps(info), flush=True)
            out.append(info)
            break            # context is likely poisoned
        print(json.dumps(info), flush=True)
        out.append(info)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/diag_conv_%s.json' % (names[0] if len(names) == 1 else 'multi_%d' % os.getpid()), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
