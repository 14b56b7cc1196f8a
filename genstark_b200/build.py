"""Builds libgenstark_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgenstark_b200.so')
SOURCES = ['api.cu']
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-shared']


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, '..', 'include', 'genstark_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    # host-only code goes through the host compiler directly (better code for the sequential trace loop)
    host_obj = os.path.join(HERE, 'host.o')
    cxx = os.environ.get('CXX', 'g++')
    hcmd = [cxx, '-O3', '-std=c++17', '-fPIC', '-pthread', '-c', os.path.join(CSRC, 'host.cpp'), '-o', host_obj]
    cmd = [nvcc, *NVCC_FLAGS, '-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES] + [host_obj]
    if verbose:
        cmd.insert(1, '-Xptxas'); cmd.insert(2, '-v')
        print(' '.join(hcmd), file=sys.stderr)
        print(' '.join(cmd), file=sys.stderr)
    subprocess.run(hcmd, check=True)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
    print(LIB)
