"""Build librgnn.so (hand-written sm_100a CUDA kernels + the C-ABI of include/rgnn.h) in-tree with nvcc.

    python -m graph_neural_network_for_radar_perception_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting csrc/librgnn.so travels to the GPU box with the repo
snapshot.  Objects are cached under csrc/build/ keyed by source mtime.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(CSRC, 'build')
LIB = os.path.join(CSRC, 'librgnn.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; librgnn.so cannot be built (there is no CPU fallback)')


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime() -> float:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(os.path.dirname(HERE), 'include', 'rgnn.h'))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    hdr_m = _deps_mtime()
    objs, rebuilt = [], False
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src[:-3] + '.o')
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            cmd = [nvcc] + NVCC_FLAGS + ['-c', s, '-o', o]
            if verbose:
                print('[rgnn build]', ' '.join(cmd), flush=True)
            subprocess.check_call(cmd)
            rebuilt = True
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        if verbose:
            print('[rgnn build]', ' '.join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
