"""Builds liborbx_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc. No JIT cache: the .so sits next
to the sources so that it travels with a repository snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIB_DIR, 'liborbx_b200.so')
SOURCES = ['orbx_extract.cu', 'orbx_match.cu', 'orbx_guided.cu', 'orbx_bow.cu', 'orbx_api.cu']
DEPS = SOURCES + ['orbx_internal.cuh', 'orbx_sort.cuh', 'orbx_quadtree.cuh', 'orbx_strip.cuh', 'orbx_describe.cuh', 'orb_pattern.inc', os.path.join('..', '..', 'include', 'orbx.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '--fmad=false',                                   # float ops of the reference are never contracted
    '-Xcompiler', '-fPIC,-ffp-contract=off,-O2',
    '-shared',
]


def nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed building liborbx_b200.so')
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
