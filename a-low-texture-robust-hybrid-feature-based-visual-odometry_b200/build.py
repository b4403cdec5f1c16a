"""Build recipe for libhvofront.so (explicit nvcc, sm_100a only, built in-tree so it travels with gpurun)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libhvofront.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-fmad=false',                      # bit-exact float semantics: no FMA contraction anywhere
    '-Xcompiler', '-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function',
    '-shared', '-cudart', 'static',
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '..', 'include', '*'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, extra=()):
    """out / extra: tuning aid, builds a variant (e.g. extra=['-DHVO_FAST_THREADS=64']) next to the product library."""
    if out is None and not force and not needs_build():
        return OUT
    out = out or OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (['-Xptxas', '-v'] if verbose else []) + ['-o', out] + sources()
    print(' '.join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return out


if __name__ == '__main__':
    defs = [a for a in sys.argv[1:] if a.startswith('-D')]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith('--out=')]
    build(force='--force' in sys.argv, verbose='-v' in sys.argv, out=outs[0] if outs else None, extra=defs)
