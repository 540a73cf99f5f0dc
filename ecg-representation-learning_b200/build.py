"""
In-tree build of the C-ABI kernel library `libecgvit_b200.so` for sm_100a.

`nvcc` cross-compiles here without a GPU; the resulting .so travels to the GPU box with the repo snapshot.
cudart is linked statically and the only driver entry point used (cuTensorMapEncodeTiled) is resolved at run
time, so the library loads on a machine without a GPU or libcuda (the CPU test tier checks its exports).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
BUILD_DIR = os.path.join(PKG_DIR, 'build')
LIB_PATH = os.path.join(PKG_DIR, 'libecgvit_b200.so')
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), 'include')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC',
]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; cannot build libecgvit_b200.so')
    return exe


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    hs.append(os.path.join(INCLUDE_DIR, 'ecgvit_b200.h'))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link libecgvit_b200.so; no-op when up to date."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = _headers()
    jobs = []
    objs = []
    for src in _sources():
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed: {" ".join(cmd)}\n{r.stdout}\n{r.stderr}')
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(LIB_PATH, objs):
        run([nvcc, '-shared', '-o', LIB_PATH] + objs + ['-cudart', 'static'])
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
