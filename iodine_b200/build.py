"""Build libiodine_b200.so in-tree with nvcc for sm_100a (no torch involved).

    python -m iodine_b200.build [--force]

The shared library is a plain C-ABI object (include/iodine_b200.h); it travels to the GPU
box with the repo snapshot (git-ignored, not gpurun-ignored).
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, 'lib')
LIB = os.path.join(OUT_DIR, 'libiodine_b200.so')
SOURCES = ['plan.cu', 'conv_f32.cu', 'mixture.cu', 'head.cu', 'conv_tc.cu', 'refine_tc.cu', 'ari.cu', 'train.cu', 'wgrad_tc.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for fn in sorted(os.listdir(root)):
            if fn.endswith(('.cu', '.cuh', '.h')):
                with open(os.path.join(root, fn), 'rb') as f:
                    h.update(fn.encode())
                    h.update(f.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OUT_DIR, src.replace('.cu', '.o'))
    cmd = [NVCC] + FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, 'build.stamp')
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError('nvcc not found at %s and no prebuilt %s' % (NVCC, LIB))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        res = list(ex.map(_compile, SOURCES))
    objs = [o for o, _ in res]
    if verbose:
        for _, err in res:
            if err.strip():
                print(err)
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                '-Xcompiler', '-fPIC']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    with open(stamp, 'w') as f:
        f.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
