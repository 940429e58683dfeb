"""In-tree build of libdbsr_b200.so (nvcc, sm_100a only).

    python -m deep_rawburst_sr_b200.build [--force]

The shared object is written next to the sources (deep_rawburst_sr_b200/csrc/libdbsr_b200.so), is git-ignored,
and travels to the GPU box with the repo snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(CSRC, 'libdbsr_b200.so')
SOURCES = ['misc.cu', 'conv_direct.cu', 'corr.cu', 'fusion.cu', 'metrics.cu', 'camera.cu', 'conv_tc.cu', 'resblock_tc.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, 'common.cuh'), os.path.join(CSRC, 'tma.cuh'), os.path.join(CSRC, 'tcgen05.cuh'),
                                                       os.path.join(HERE, '..', 'include', 'dbsr_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(CSRC, src[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        for obj, log in ex.map(compile_one, SOURCES):
            objs.append(obj)
            if verbose:
                sys.stderr.write(log)
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-cudart', 'static']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    path = build_library(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(path)
