"""CPU: the C-ABI shared library loads and exports every symbol include/yolo2_b200.h declares
(no compute calls without a GPU)."""
import os
import re

from tensorflow_yolo2_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'yolo2_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(y2_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_every_declared_symbol():
    from tensorflow_yolo2_b200 import build
    build.build()
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), 'missing export: ' + s
    assert sorted(_lib.exported_symbols()) == syms, 'ctypes binding table out of sync with the header'
    assert lib.y2_version() == 100
    # pure host helpers (no GPU needed)
    assert lib.y2_conv_cin_padded(3) == 8 and lib.y2_conv_cin_padded(64) == 64
    assert lib.y2_conv_packed_weight_elems(3, 3, 32) == 32 * 80
    assert lib.y2_conv_packed_weight_elems(1, 1024, 125) == 128 * 1024


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from tensorflow_yolo2_b200 import ops
    with pytest.raises(_lib.Y2Error):
        ops._p(torch.zeros(4))


def test_conv1_fused_byte_conversion_identity():
    """conv1_fused.cu converts a byte with ONE fma, bf16_rn(fma(v, 2/255, -1)); the preprocessing kernels (and the
    reference, pascal_voc.py:62-64) compute (v/255)*2-1 in float32.  The two agree after the bf16 rounding for every
    byte value -- that identity is what makes the fused first layer's input bit-identical."""
    import numpy as np
    import torch
    v = np.arange(256, dtype=np.float32)
    ref = ((v / np.float32(255.0)) * np.float32(2.0) - np.float32(1.0)).astype(np.float32)
    k = np.float32(2.0 / 255.0)
    fma = (v.astype(np.float64) * np.float64(k) - 1.0).astype(np.float32)      # exact product, one rounding
    a = torch.tensor(ref).to(torch.bfloat16).view(torch.int16)
    b = torch.tensor(fma).to(torch.bfloat16).view(torch.int16)
    assert torch.equal(a, b)
