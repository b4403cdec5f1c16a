"""Build recipe for libhvofront.so (explicit nvcc, sm_100a only, built in-tree so it travels with gpurun).

Every csrc/*.cu is compiled to its own object (in parallel, only when it or a header changed) and the objects are linked
into one shared library; there is no relocatable device code, so the objects are independent."""
import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libhvofront.so')
OBJ = os.path.join(HERE, 'build')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-fmad=false',                      # bit-exact float semantics: no FMA contraction anywhere
    '-Xcompiler', '-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function',
]
LINK_FLAGS = ['-shared', '-cudart', 'static']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def headers():
    return sorted(glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '..', 'include', '*')))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in sources() + headers())


def _obj_path(src, extra):
    tag = hashlib.sha1(' '.join(extra).encode()).hexdigest()[:8] if extra else 'std'
    return os.path.join(OBJ, os.path.basename(src)[:-3] + '.' + tag + '.o')


def build(force=False, verbose=False, out=None, extra=()):
    """out / extra: tuning aid, builds a variant (e.g. extra=['-DHVO_FAST_THREADS=64']) next to the product library."""
    if out is None and not force and not needs_build():
        return OUT
    out = out or OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in headers())
    extra = list(extra)

    def compile_one(src):
        obj = _obj_path(src, extra)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, src]
        print(' '.join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    cmd = [nvcc] + NVCC_FLAGS + LINK_FLAGS + ['-o', out] + objs
    print(' '.join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return out


if __name__ == '__main__':
    defs = [a for a in sys.argv[1:] if a.startswith('-D')]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith('--out=')]
    build(force='--force' in sys.argv, verbose='-v' in sys.argv, out=outs[0] if outs else None, extra=defs)
