"""Timing + check of the two big-graph (V = 3100) kernels outside the hop GEMM: the support gradient `gwn_dadj_big`
(split-K; GWN_DADJ_SPLITS=n forces the number of K parts, read once per process) and the sparse hop `gwn_hop_ell`.
Shapes of BASELINE config 3 (batch 64): slabs = 64 * L for L = 12, 7, 3."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import numpy as np
import torch

from multimodal_outage_b200.supports import double_transition
from bench import synthetic_knn_graph
from multimodal_outage_b200 import _lib, ops

lib = _lib.lib()
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
torch.manual_seed(0)
V = 3100


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print('GWN_DADJ_SPLITS =', os.environ.get('GWN_DADJ_SPLITS', '(auto)'))
for L in (12, 7, 3):
    slabs = 64 * L
    x = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
    g = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
    dA = torch.zeros(V, V, device='cuda')
    _lib.check(lib.gwn_dadj_big(x.data_ptr(), g.data_ptr(), dA.data_ptr(), slabs, V, st()), 'dadj')
    if L == 3:
        ref = torch.einsum('svc,swc->vw', x.float(), g.float())
        print(f'  dadj check (L=3): rel {rel(dA, ref):.2e}')
    us = timed(lambda: _lib.check(lib.gwn_dadj_big(x.data_ptr(), g.data_ptr(), dA.data_ptr(), slabs, V, st()), 'dadj'))
    fl = 2.0 * V * V * slabs * 32
    print(f'dadj_big slabs={slabs}: {us:8.1f} us  {fl / us * 1e-6:7.1f} TFLOP/s')

# dense hop (adaptive support): y = Aop x over `slabs` slabs
Ad = torch.softmax(torch.randn(V, V, device='cuda'), dim=1).contiguous()
img = ops.support_images([Ad])
for L in (12, 7, 3):
    slabs = 64 * L
    x = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
    y = torch.empty_like(x)
    for which in range(2):
        f = lambda: _lib.check(lib.gwn_hop_big(img.data_ptr(), 1, 0, which, x.data_ptr(), y.data_ptr(), None, slabs, V, st()), 'hop')  # noqa: E731
        us = timed(f)
        if L == 3:
            ref = torch.einsum('vw,svc->swc' if which == 0 else 'wv,svc->swc', Ad.to(torch.bfloat16).float(), x.float())
            print(f'  hop check which={which}: rel {rel(y, ref):.2e}')
        print(f'hop_big slabs={slabs} which={which}: {us:8.1f} us  {2.0 * V * V * slabs * 32 / us * 1e-6:7.1f} TFLOP/s')

A = torch.tensor(np.asarray(double_transition(synthetic_knn_graph(V))[0]), dtype=torch.float32, device='cuda').contiguous()
assert ops.register_sparse_support(A)
e = ops._ELL_REGISTRY[A.data_ptr()]
print('ELL width', e['width'], 'mean degree', float((A != 0).sum(dim=1).float().mean()))
for L in (12, 7, 3):
    slabs = 64 * L
    x = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
    y = torch.empty_like(x)
    for which in range(2):
        f = lambda: _lib.check(lib.gwn_hop_ell(e['idx'][which].data_ptr(), e['val'][which].data_ptr(), e['width'], x.data_ptr(),  # noqa: E731
                                               y.data_ptr(), None, slabs, V, st()), 'ell')
        us = timed(f)
        if L == 3:
            ref = torch.einsum('vw,svc->swc' if which == 0 else 'wv,svc->swc', A, x.float())
            print(f'  ell check which={which}: rel {rel(y, ref):.2e}')
        print(f'hop_ell slabs={slabs} which={which}: {us:8.1f} us  {2 * x.numel() * 2 / us * 1e-3:7.1f} GB/s (x in + y out)')
