"""First light + timing for the TMA-fed tcgen05 GEMM (csrc/tma_gemm.cuh) and the big-V diffusion hop."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch

from multimodal_outage_b200 import _lib

lib = _lib.lib()
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
torch.manual_seed(0)
dev = 'cuda'


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def gemm_case(M, N, K, a_mode, b_mode, bn, splits=1):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    pad8 = lambda x: (x + 7) // 8 * 8  # noqa: E731
    if a_mode == 0:
        lda = pad8(K); Ag = torch.zeros(M, lda, device=dev, dtype=torch.bfloat16); Ag[:, :K] = A
    else:
        lda = pad8(M); Ag = torch.zeros(K, lda, device=dev, dtype=torch.bfloat16); Ag[:, :M] = A.t()
    if b_mode == 0:
        ldb = pad8(K); Bg = torch.zeros(N, ldb, device=dev, dtype=torch.bfloat16); Bg[:, :K] = B
    elif b_mode == 1:
        ldb = pad8(N); Bg = torch.zeros(K, ldb, device=dev, dtype=torch.bfloat16); Bg[:, :N] = B.t()
    else:
        assert N % 32 == 0
        ldb = 32; Bg = B.reshape(N // 32, 32, K).permute(0, 2, 1).contiguous()      # [slab][K][32]
    Cc = torch.zeros(M, N, device=dev)
    rc = lib.gwn_gemm_test(Ag.data_ptr(), Bg.data_ptr(), Cc.data_ptr(), M, N, K, a_mode, b_mode, lda, ldb, bn, splits, st())
    if rc != 0:
        print(f'  gemm M={M} N={N} K={K} a{a_mode} b{b_mode} bn={bn}: ERROR {lib.gwn_last_error().decode()}')
        return None
    torch.cuda.synchronize()
    e = rel(Cc, ref)
    print(f'  gemm M={M} N={N} K={K} a{a_mode} b{b_mode} bn={bn} splits={splits}: rel err {e:.3e}', flush=True)
    return e


print('== staging modes')
for (am, bm) in ((0, 0), (1, 0), (0, 1), (0, 2), (1, 2), (1, 1)):
    for (M, N, K, bn) in ((128, 256, 64, 256), (256, 512, 192, 256), (300, 320, 200, 64), (3100, 512, 3100, 256)):
        gemm_case(M, N, K, am, bm, bn)
gemm_case(256, 256, 4096, 0, 0, 256, splits=8)
gemm_case(256, 512, 34304, 1, 1, 256, splits=37)

print('== big-V hop')
for V, slabs in ((200, 16), (3100, 24), (3100, 768)):
    Vp = (V + 7) // 8 * 8
    sups = [torch.softmax(torch.randn(V, V, device=dev), dim=1) for _ in range(2)]
    nbytes = lib.gwn_support_images_bytes(V, 2)
    img = torch.empty(nbytes // 2, device=dev, dtype=torch.bfloat16)
    ptrs = (C.c_void_p * 2)(*[s.data_ptr() for s in sups])
    _lib.check(lib.gwn_support_images_prep(ptrs, 2, V, img.data_ptr(), st()), 'prep')
    x = torch.randn(slabs, V, 32, device=dev).to(torch.bfloat16)
    add = torch.randn(slabs, V, 32, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    for s in range(2):
        Ab = sups[s].to(torch.bfloat16).float()
        for which in range(2):
            _lib.check(lib.gwn_hop_big(img.data_ptr(), 2, s, which, x.data_ptr(), y.data_ptr(), None, slabs, V, st()), 'hop')
            torch.cuda.synchronize()
            ref = torch.einsum('vw,svc->swc' if which == 0 else 'wv,svc->swc', Ab, x.float())
            e1 = rel(y, ref)
            _lib.check(lib.gwn_hop_big(img.data_ptr(), 2, s, which, x.data_ptr(), y.data_ptr(), add.data_ptr(), slabs, V, st()), 'hop')
            torch.cuda.synchronize()
            e2 = rel(y, ref + add.float())
            print(f'  V={V} slabs={slabs} support {s} which {which}: rel err {e1:.3e} (with add {e2:.3e})', flush=True)
    if slabs >= 768:
        # timing: rotate over 3 distinct activation buffers (152 MB each > L2)
        xs = [torch.randn(slabs, V, 32, device=dev).to(torch.bfloat16) for _ in range(3)]
        ys = [torch.empty_like(x) for _ in range(3)]
        for i in range(3):
            lib.gwn_hop_big(img.data_ptr(), 2, 0, 0, xs[i].data_ptr(), ys[i].data_ptr(), None, slabs, V, st())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 12
        e0.record()
        for i in range(iters):
            lib.gwn_hop_big(img.data_ptr(), 2, 0, 0, xs[i % 3].data_ptr(), ys[i % 3].data_ptr(), None, slabs, V, st())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * slabs * 32 * V * V
        print(f'  hop_big V={V} slabs={slabs}: {ms:.3f} ms/hop, {fl / ms / 1e9:.1f} TFLOP/s algorithmic', flush=True)
        # cuBLAS reference for the same contraction
        Ab = sups[0].to(torch.bfloat16)
        xf = xs[0].permute(1, 0, 2).reshape(V, slabs * 32).contiguous()
        for _ in range(3): torch.matmul(Ab.t(), xf)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters): torch.matmul(Ab.t(), xf)
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / iters
        print(f'  cuBLAS same GEMM (pre-transposed activations): {ms2:.3f} ms, {fl / ms2 / 1e9:.1f} TFLOP/s', flush=True)
