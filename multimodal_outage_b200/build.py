"""Builds libgwn.so (hand-written sm_100a CUDA, plain C ABI) in-tree with nvcc.

No torch C++ extension: the library has no torch types in its signatures
(include/gwn.h), so it is compiled with plain ``nvcc -shared`` and bound via ctypes.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgwn.so')
STAMP = os.path.join(HERE, 'libgwn.so.stamp')
SOURCES = ['api.cu', 'adp.cu', 'layer.cu', 'head.cu', 'peer.cu', 'tc_hops.cu', 'tc_wgrad.cu', 'tc_gemm.cu', 'tma_gemm.cu', 'head_tc.cu', 'gcn_fused.cu', 'gcn_fused_bwd.cu', 'gcn_fused_bwd_t.cu', 'start_tc.cu', 'pack.cu', 'gcn_fused_t.cu', 'loss.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']
TRACE = os.environ.get('GWN_TRACE') == '1'    # debug build: clock64 timeline hooks read GWN_*_TRACE (scripts/gpu_*_trace.py)
if TRACE:                                     # goes to its own library (libgwn_trace.so, loaded with GWN_LIB=...) beside the product
    NVCC_FLAGS.append('-DGWN_TRACE')
    LIB = os.path.join(HERE, 'libgwn_trace.so')
    STAMP = LIB + '.stamp'


VARIANT = os.environ.get('GWN_VARIANT')        # experiment builds: GWN_VARIANT=name GWN_DEFINES="-DX=1 ..." -> libgwn_<name>.so
if VARIANT:
    NVCC_FLAGS += os.environ.get('GWN_DEFINES', '').split()
    LIB = os.path.join(HERE, f'libgwn_{VARIANT}.so')
    STAMP = LIB + '.stamp'


def _fingerprint() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for fn in sorted(os.listdir(root)):
            if fn.endswith(('.cu', '.cuh', '.h')):
                h.update(fn.encode())
                h.update(open(os.path.join(root, fn), 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == fp:
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    bdir = os.path.join(HERE, 'build', VARIANT or ('trace' if TRACE else ''))
    os.makedirs(bdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, '-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f'== {src}\n{out}')
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
    link = [nvcc, '-shared', '-o', LIB, *objs, '-lcudart', '-ldl']
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}')
    with open(os.path.join(bdir, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    with open(STAMP, 'w') as f:
        f.write(fp)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
