"""Build libyolo2_b200.so (all CUDA kernels + the C ABI) in-tree with nvcc for sm_100a.

    python -m tensorflow_yolo2_b200.build [--force] [--verbose]

Output: tensorflow_yolo2_b200/lib/libyolo2_b200.so  (git-ignored, travels with gpurun snapshots).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libyolo2_b200.so')
SOURCES = ['api.cu', 'elementwise.cu', 'conv_simt.cu', 'conv_tcgen05.cu', 'conv_streamk_tcgen05.cu', 'conv1_fused.cu', 'decode.cu', 'nms.cu', 'detect_fused.cu', 'detect_split.cu', 'loss.cu', 'region_loss.cu', 'backward.cu',
           'conv_wgrad_tcgen05.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _newest_source_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(HERE, '..', 'include', 'yolo2_b200.h'))
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('--- %s ---\n%s\n' % (src, out))
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
