"""Builds libfibb200.so in-tree with nvcc for sm_100a:  python -m fib_tf_b200.build [-v]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'fib_capi.cu')
OUT = os.path.join(HERE, 'libfibb200.so')
# -fmad=false: multiply-adds are fused only where the source says so (fmaf / vfma / fma.rn.f32x2).  That
# makes every kernel that shares a cell function -- one step per launch, two steps per launch, the
# persistent on-chip kernel, the scalar and the packed (f32x2) flavours -- round identically, so they can
# be (and are) tested BIT-identical to each other; left to the compiler, contraction differs per kernel.
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-fmad=false',
              '-Xcompiler', '-fPIC', '-shared']


def sources():
    d = os.path.join(HERE, 'csrc')
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [
        os.path.join(os.path.dirname(HERE), 'include', 'fib_b200.h')]


def up_to_date():
    return os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in sources())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUT, SRC, '-ldl']
    print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    build(force='-f' in sys.argv, verbose='-v' in sys.argv)
