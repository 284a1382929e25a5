"""torch.library custom ops (`torch.ops.gwn.*`) over the libgwn C ABI, plus the
autograd Functions that chain them into the Graph WaveNet block.

Every op calls hand-written sm_100a kernels through ctypes with raw device pointers
and the current CUDA stream.  There is no CPU / eager fallback: non-CUDA tensors or a
non-B200 device raise.

Internal activation layout is channels-last ``[N, L, V, 32]`` (see include/gwn.h).
Reference lines replaced are cited per op (``/root/reference/models/graph_wavenet.py``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import (GWN_BF16, GWN_F32, HeadBwdArgs, HeadCfg, HeadFwdArgs, HeadTcBwdArgs, HeadTcFwdArgs, LayerBwdArgs,
                   LayerCfg, LayerFwdArgs, PackCfg, PackPtrs, UnpackPtrs, check, lib)

CH = 32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return GWN_F32
    if dtype == torch.bfloat16:
        return GWN_BF16
    raise TypeError(f'activation dtype must be float32 or bfloat16, got {dtype}')


def _p(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: Tensor, dtype=None, name='tensor') -> Tensor:
    if not t.is_cuda:
        raise _lib.GwnError(f'{name}: libgwn ops run on CUDA (B200) tensors only - there is no CPU fallback')
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f'{name}: expected {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name}: must be contiguous')
    _lib.require_b200(t.device.index)
    return t


# =========================================================================== adaptive adjacency (:202)
@torch.library.custom_op('gwn::adp_fwd', mutates_args=())
def adp_fwd(e1: Tensor, e2: Tensor) -> Tensor:
    _req(e1, torch.float32, 'nodevec1'); _req(e2, torch.float32, 'nodevec2')
    V, R = e1.shape
    adp = torch.empty((V, V), device=e1.device, dtype=torch.float32)
    with torch.cuda.device(e1.device):
        check(lib().gwn_adp_fwd(_p(e1), _p(e2), _p(adp), None, V, R, _stream()), 'gwn_adp_fwd')
    return adp


@adp_fwd.register_fake
def _(e1, e2):
    return e1.new_empty((e1.shape[0], e1.shape[0]))


@torch.library.custom_op('gwn::adp_bwd', mutates_args=())
def adp_bwd(e1: Tensor, e2: Tensor, adp: Tensor, d_adp: Tensor) -> Tuple[Tensor, Tensor]:
    _req(d_adp, torch.float32, 'd_adp')
    V, R = e1.shape
    de1, de2 = torch.empty_like(e1), torch.empty_like(e2)
    ws = torch.empty((V, V), device=e1.device, dtype=torch.float32)
    with torch.cuda.device(e1.device):
        check(lib().gwn_adp_bwd(_p(e1), _p(e2), _p(adp), _p(d_adp), _p(de1), _p(de2), _p(ws), V, R, _stream()),
              'gwn_adp_bwd')
    return de1, de2


@adp_bwd.register_fake
def _(e1, e2, adp, d_adp):
    return torch.empty_like(e1), torch.empty_like(e2)


@torch.library.custom_op('gwn::adp_fwd_pair', mutates_args=())
def adp_fwd_pair(e1: Tensor, e2: Tensor) -> Tensor:
    """The adaptive adjacency as a pair [2,V,V] of identical copies (include/gwn.h: gwn_adp_fwd_pair)."""
    _req(e1, torch.float32, 'nodevec1'); _req(e2, torch.float32, 'nodevec2')
    V, R = e1.shape
    pair = torch.empty((2, V, V), device=e1.device, dtype=torch.float32)
    with torch.cuda.device(e1.device):
        check(lib().gwn_adp_fwd_pair(_p(e1), _p(e2), _p(pair), V, R, _stream()), 'gwn_adp_fwd_pair')
    return pair


@adp_fwd_pair.register_fake
def _(e1, e2):
    return e1.new_empty((2, e1.shape[0], e1.shape[0]))


@torch.library.custom_op('gwn::adp_pair_bwd', mutates_args=())
def adp_pair_bwd(e1: Tensor, e2: Tensor, pair: Tensor, d_pair: Tensor) -> Tuple[Tensor, Tensor]:
    """d_pair = (d0, Q) -> d_adp = d0 + A^T Q + Q A^T (fp32) -> softmax / relu / rank-R backward."""
    _req(d_pair, torch.float32, 'd_pair')
    V, R = e1.shape
    de1, de2 = torch.empty_like(e1), torch.empty_like(e2)
    ws = torch.empty((2, V, V), device=e1.device, dtype=torch.float32)
    with torch.cuda.device(e1.device):
        check(lib().gwn_adp_pair_bwd(_p(e1), _p(e2), _p(pair), _p(d_pair), _p(de1), _p(de2), _p(ws), V, R, _stream()),
              'gwn_adp_pair_bwd')
    return de1, de2


@adp_pair_bwd.register_fake
def _(e1, e2, pair, d_pair):
    return torch.empty_like(e1), torch.empty_like(e2)


class AdaptiveAdjacency(torch.autograd.Function):
    """softmax(relu(E1 @ E2), dim=1), warp-per-row fused forward and backward.

    ``pair=True`` returns ``[2,V,V]`` (two identical copies).  A layer that receives the pair as a support may return
    the gradient in factored form ``(d0, Q)`` - the part reaching the adjacency through its A*A hop
    (graph_wavenet.py:91-93) is linear in Q = sum (z W_{A^2})[v] . dh[w] - and the backward here completes it once
    per step in fp32: ``d_adp = d0 + A^T Q + Q A^T`` (csrc/gcn_fused_bwd_t.cu)."""

    @staticmethod
    def forward(ctx, e1, e2, pair=False):
        e1c, e2c = e1.contiguous(), e2.contiguous()
        ctx.pair = bool(pair)
        adp = adp_fwd_pair(e1c, e2c) if ctx.pair else adp_fwd(e1c, e2c)
        ctx.save_for_backward(e1c, e2c, adp)
        return adp

    @staticmethod
    def backward(ctx, d_adp):
        e1, e2, adp = ctx.saved_tensors
        if ctx.pair:
            return (*adp_pair_bwd(e1, e2, adp, d_adp.contiguous()), None)
        return (*adp_bwd(e1, e2, adp, d_adp.contiguous()), None)


# =========================================================================== start conv (:191-196)
@torch.library.custom_op('gwn::start_fwd', mutates_args=())
def start_fwd(x: Tensor, w: Tensor, b: Tensor, L0: int, bf16: bool) -> Tensor:
    _req(x, torch.float32, 'input'); _req(w, torch.float32, 'start_conv.weight'); _req(b, torch.float32)
    N, Cin, V, T = x.shape
    dt = torch.bfloat16 if bf16 else torch.float32
    u0 = torch.empty((N, L0, V, CH), device=x.device, dtype=dt)
    with torch.cuda.device(x.device):
        check(lib().gwn_start_fwd(_p(x), _p(w), _p(b), _p(u0), _code(dt), N, Cin, V, T, L0, _stream()),
              'gwn_start_fwd')
    return u0


@start_fwd.register_fake
def _(x, w, b, L0, bf16):
    return x.new_empty((x.shape[0], L0, x.shape[2], CH), dtype=torch.bfloat16 if bf16 else torch.float32)


@torch.library.custom_op('gwn::start_bwd', mutates_args=())
def start_bwd(x: Tensor, w: Tensor, du0: Tensor, need_dx: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _req(du0, None, 'du0')
    N, Cin, V, T = x.shape
    L0 = du0.shape[1]
    dw = torch.empty((CH, Cin), device=x.device, dtype=torch.float32)
    db = torch.empty((CH,), device=x.device, dtype=torch.float32)
    dx = torch.empty_like(x) if need_dx else torch.empty((0,), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(lib().gwn_start_bwd(_p(x), _p(w), _p(du0), _code(du0.dtype), _p(dw), _p(db),
                                  _p(dx) if need_dx else None, N, Cin, V, T, L0, _stream()), 'gwn_start_bwd')
    return dw, db, dx


@start_bwd.register_fake
def _(x, w, du0, need_dx):
    return (x.new_empty((CH, x.shape[1])), x.new_empty((CH,)),
            torch.empty_like(x) if need_dx else x.new_empty((0,)))


@torch.library.custom_op('gwn::start_fwd_tc', mutates_args=())
def start_fwd_tc(x: Tensor, w: Tensor, b: Tensor, L0: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Wide-input start conv on tensor cores: returns (u0 bf16 [N,L0,V,32], xcl bf16 [N,L0,V,Cin], ws_w)."""
    _req(x, torch.float32, 'input'); _req(w, torch.float32, 'start_conv.weight'); _req(b, torch.float32)
    N, Cin, V, T = x.shape
    dev = x.device
    u0 = torch.empty((N, L0, V, CH), device=dev, dtype=torch.bfloat16)
    xcl = torch.empty((N, L0, V, Cin), device=dev, dtype=torch.bfloat16)
    ws_w = torch.empty((2 * CH * Cin,), device=dev, dtype=torch.bfloat16)
    with torch.cuda.device(dev):
        check(lib().gwn_start_fwd_tc(_p(x), _p(w), _p(b), _p(xcl), _p(u0), _p(ws_w), N, Cin, V, T, L0, _stream()),
              'gwn_start_fwd_tc')
    return u0, xcl, ws_w


@start_fwd_tc.register_fake
def _(x, w, b, L0):
    N, Cin, V, T = x.shape
    return (x.new_empty((N, L0, V, CH), dtype=torch.bfloat16), x.new_empty((N, L0, V, Cin), dtype=torch.bfloat16),
            x.new_empty((2 * CH * Cin,), dtype=torch.bfloat16))


@torch.library.custom_op('gwn::start_bwd_tc', mutates_args=())
def start_bwd_tc(xcl: Tensor, ws_w: Tensor, du0: Tensor, T: int, need_dx: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _req(du0, torch.bfloat16, 'du0')
    N, L0, V, Cin = xcl.shape
    dev = xcl.device
    dw = torch.empty((CH, Cin), device=dev, dtype=torch.float32)
    db = torch.empty((CH,), device=dev, dtype=torch.float32)
    dx = torch.empty((N, Cin, V, T) if need_dx else (0,), device=dev, dtype=torch.float32)
    ws_dx = torch.empty((N * L0 * V * Cin,) if need_dx else (0,), device=dev, dtype=torch.bfloat16)
    with torch.cuda.device(dev):
        check(lib().gwn_start_bwd_tc(_p(xcl), _p(du0), _p(ws_w), _p(dw), _p(db), _p(dx) if need_dx else None,
                                     _p(ws_dx) if need_dx else None, N, Cin, V, T, L0, _stream()), 'gwn_start_bwd_tc')
    return dw, db, dx


@start_bwd_tc.register_fake
def _(xcl, ws_w, du0, T, need_dx):
    N, L0, V, Cin = xcl.shape
    return (xcl.new_empty((CH, Cin), dtype=torch.float32), xcl.new_empty((CH,), dtype=torch.float32),
            xcl.new_empty((N, Cin, V, T) if need_dx else (0,), dtype=torch.float32))


class StartConv(torch.autograd.Function):
    """1x1 conv Cin->32 fused with the left zero-pad and the NCHW -> channels-last change."""

    @staticmethod
    def forward(ctx, x, w, b, L0, bf16):
        xc = x.contiguous().float()
        w2 = w.reshape(CH, -1).contiguous()
        ctx.wshape = w.shape
        ctx.tc = bool(bf16) and bool(lib().gwn_start_tc_supported(xc.shape[1]))
        if ctx.tc:      # wide inputs (config 4): transpose to channels-last bf16 once, then a tensor-core GEMM
            u0, xcl, ws_w = start_fwd_tc(xc, w2, b.contiguous(), L0)
            ctx.save_for_backward(xcl, ws_w)
            ctx.T = xc.shape[3]
            return u0
        u0 = start_fwd(xc, w2, b.contiguous(), L0, bf16)
        ctx.save_for_backward(xc, w2)
        return u0

    @staticmethod
    def backward(ctx, du0):
        need_dx = ctx.needs_input_grad[0]
        if ctx.tc:
            xcl, ws_w = ctx.saved_tensors
            dw, db, dx = start_bwd_tc(xcl, ws_w, du0.contiguous(), ctx.T, need_dx)
            return (dx if need_dx else None), dw.reshape(ctx.wshape), db, None, None
        xc, w2 = ctx.saved_tensors
        dw, db, dx = start_bwd(xc, w2, du0.contiguous(), need_dx)
        return (dx if need_dx else None), dw.reshape(ctx.wshape), db, None, None


# =========================================================================== BatchNorm fold (:167,250)
@torch.library.custom_op('gwn::bn_fold', mutates_args=('running_mean', 'running_var'))
def bn_fold(stats: Optional[Tensor], count: float, gamma: Tensor, beta: Tensor, running_mean: Tensor,
            running_var: Tensor, momentum: float, eps: float, training: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    _req(gamma, torch.float32, 'bn.weight')
    scale, shift, mean, rstd = (torch.empty_like(gamma) for _ in range(4))
    with torch.cuda.device(gamma.device):
        check(lib().gwn_bn_fold(_p(stats), float(count), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
                                float(momentum), float(eps), int(training), _p(scale), _p(shift), _p(mean),
                                _p(rstd), _stream()), 'gwn_bn_fold')
    return scale, shift, mean, rstd


@bn_fold.register_fake
def _(stats, count, gamma, beta, running_mean, running_var, momentum, eps, training):
    return tuple(torch.empty_like(gamma) for _ in range(4))


@torch.library.custom_op('gwn::bn_bwd', mutates_args=())
def bn_bwd(dx: Tensor, u: Tensor, dx_stats: Tensor, count: float, gamma: Tensor, mean: Tensor, rstd: Tensor,
           training: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _req(dx, None, 'dx'); _req(u, None, 'u')
    du = torch.empty_like(u)
    dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
    rows = u.numel() // CH
    with torch.cuda.device(u.device):
        check(lib().gwn_bn_bwd(_p(dx), _code(dx.dtype), _p(u), _code(u.dtype), _p(dx_stats), float(count), _p(gamma), _p(mean),
                               _p(rstd), int(training), _p(du), _p(dgamma), _p(dbeta), rows, _stream()),
              'gwn_bn_bwd')
    return du, dgamma, dbeta


@bn_bwd.register_fake
def _(dx, u, dx_stats, count, gamma, mean, rstd, training):
    return torch.empty_like(u), torch.empty_like(gamma), torch.empty_like(gamma)


# =========================================================================== sparse fixed supports (V > 80)
# County adjacency graphs have a handful of neighbours per node, and asym_adj keeps that pattern: at V = 3100 a dense hop
# multiplies 99.7 % exact zeros.  A support registered here is applied by the big-graph bf16 path with a gather kernel over
# ELL rows (include/gwn.h: gwn_ell) - same sums, zero terms skipped.  Keyed by the support tensor's storage; an entry is
# used only while that very tensor is alive and unmodified.
_ELL_REGISTRY: dict = {}
ELL_MAX_WIDTH = 32


def register_sparse_support(A: Tensor, max_width: int = ELL_MAX_WIDTH) -> bool:
    """Builds ELL rows (both hop directions) of a fixed [V,V] fp32 CUDA support and registers them.  Returns False (and
    registers nothing) when some row or column has more than `max_width` non-zeros - the support then stays dense."""
    import weakref
    _req(A, torch.float32, 'support')
    V = A.shape[0]
    nz = A != 0
    width = int(max(nz.sum(dim=0).max().item(), nz.sum(dim=1).max().item(), 1))
    if width > max_width:
        _ELL_REGISTRY.pop(A.data_ptr(), None)
        return False
    # (row pitch rounded up to four entries: the hop kernel then reads a row's indices / values 16 bytes at a time)
    pitch = min(V, (width + 3) // 4 * 4)
    def rows(M):                # M[w, v] != 0 -> neighbours v of output row w, padded with -1
        order = torch.argsort((M == 0).to(torch.int8), dim=1, stable=True)[:, :pitch]       # non-zeros first, ascending v
        vals = torch.gather(M, 1, order)
        idx = torch.where(vals != 0, order, torch.full_like(order, -1)).to(torch.int32).contiguous()
        return idx, vals.contiguous()
    idx0, val0 = rows(A.t().contiguous())      # which = 0: y[w] = sum_v A[v, w] x[v]
    idx1, val1 = rows(A)                       # which = 1: y[w] = sum_v A[w, v] x[v]
    _ELL_REGISTRY[A.data_ptr()] = dict(ref=weakref.ref(A), version=A._version, idx=(idx0, idx1), val=(val0, val1), width=pitch)
    return True


def _ell_array(supports: Sequence[Tensor]):
    """ctypes `gwn_ell[MAX_SUPPORTS]` for the registered supports among `supports` (None when there is none)."""
    arr, keep = None, []
    for i, s in enumerate(supports):
        e = _ELL_REGISTRY.get(s.data_ptr())
        r = e['ref']() if e is not None else None
        if r is None or r.data_ptr() != s.data_ptr() or r.shape != s.shape or e['version'] != s._version:
            continue
        if arr is None:
            arr = (_lib.Ell * _lib.MAX_SUPPORTS)()
        arr[i].idx[0], arr[i].idx[1] = e['idx'][0].data_ptr(), e['idx'][1].data_ptr()
        arr[i].val[0], arr[i].val[1] = e['val'][0].data_ptr(), e['val'][1].data_ptr()
        arr[i].width = e['width']
        keep.append(e)
    return arr, keep


# =========================================================================== one layer (:206-250)
def _layer_cfg(N, V, Lin, Lout, Lf, taps, dilation, n_sup, order, dtype, training, has_gconv, dropout_p, seed,
               offset) -> LayerCfg:
    return LayerCfg(N=N, V=V, Lin=Lin, Lout=Lout, Lf=Lf, taps=taps, dilation=dilation, n_supports=n_sup,
                    order=order, dtype=_code(dtype), training=int(training), has_gconv=int(has_gconv),
                    dropout_p=float(dropout_p), seed=int(seed) & (2**64 - 1), offset=int(offset))


def _sup_array(supports: Sequence[Tensor]):
    arr = (C.c_void_p * _lib.MAX_SUPPORTS)()
    for i, s in enumerate(supports):
        arr[i] = s.data_ptr()
    return arr


def _layer_fwd_impl(u_prev, scale, shift, w_fg, b_fg, w_mlp, b_mlp, supports, drop_mask, rng, hop_mats, Lf, taps,
                    dilation, order, training, has_gconv, dropout_p, seed, offset, bn=None):
    _req(u_prev, None, 'u_prev'); _req(w_fg, torch.float32, 'w_fg'); _req(b_fg, torch.float32, 'b_fg')
    for s in supports:
        _req(s, torch.float32, 'support')
    N, Lin, V, _c = u_prev.shape
    Lout = Lin - dilation * (taps - 1)
    dt, dev = u_prev.dtype, u_prev.device
    n_sup = len(supports) if has_gconv else 0
    mlp_in = CH * (1 + order * n_sup)
    P = N * Lout * V
    if has_gconv:
        _req(w_mlp, torch.float32, 'w_mlp')
        if tuple(w_mlp.shape) != (mlp_in, CH):
            raise ValueError(f'w_mlp must be [{mlp_in},{CH}], got {tuple(w_mlp.shape)}')
    if tuple(w_fg.shape) != (taps * CH, 2 * CH):
        raise ValueError(f'w_fg must be [{taps * CH},{2 * CH}], got {tuple(w_fg.shape)}')
    z_last = torch.empty((N, Lf, V, CH), device=dev, dtype=dt)
    a = torch.empty((N, Lout, V, CH) if training else (0,), device=dev, dtype=dt)
    b = torch.empty((N, Lout, V, CH) if training else (0,), device=dev, dtype=dt)
    u = torch.empty((N, Lout, V, CH) if has_gconv else (0,), device=dev, dtype=dt)
    stats = torch.empty((2, CH), device=dev, dtype=torch.float64)
    ws_cat = torch.empty((P, mlp_in), device=dev, dtype=dt)
    ws_w = torch.empty((128 * 1024,), device=dev, dtype=torch.uint8) if hop_mats is not None else None
    cfg = _layer_cfg(N, V, Lin, Lout, Lf, taps, dilation, n_sup, order, dt, training, has_gconv, dropout_p, seed,
                     offset)
    args = LayerFwdArgs(u_prev=_p(u_prev), scale=_p(scale), shift=_p(shift), w_fg=_p(w_fg), b_fg=_p(b_fg),
                        w_mlp=_p(w_mlp), b_mlp=_p(b_mlp), supports=_sup_array(supports if has_gconv else []),
                        drop_mask=_p(drop_mask), rng=_p(rng), hop_mats=_p(hop_mats), ws_w=_p(ws_w),
                        a=_p(a) if training else None,
                        b=_p(b) if training else None, z_last=_p(z_last), u=_p(u) if has_gconv else None,
                        stats=_p(stats), ws_cat=_p(ws_cat))
    ell, _ell_keep = _ell_array(supports) if (has_gconv and hop_mats is not None and dt == torch.bfloat16) else (None, None)
    if ell is not None:
        args.ell = ell
    if bn is not None:
        bn_stats, count, gamma, beta, rmean, rvar, momentum, eps, mean, rstd = bn
        for t, nm in ((gamma, 'bn.weight'), (beta, 'bn.bias'), (rmean, 'running_mean'), (rvar, 'running_var')):
            _req(t, torch.float32, nm)
        args.bn_stats, args.bn_gamma, args.bn_beta = _p(bn_stats), _p(gamma), _p(beta)
        args.bn_running_mean, args.bn_running_var = _p(rmean), _p(rvar)
        args.bn_mean, args.bn_rstd = _p(mean), _p(rstd)
        args.bn_count, args.bn_momentum, args.bn_eps = float(count), float(momentum), float(eps)
    with torch.cuda.device(dev):
        check(lib().gwn_layer_fwd(C.byref(cfg), C.byref(args), _stream()), 'gwn_layer_fwd')
    return u, stats, z_last, a, b


@torch.library.custom_op('gwn::layer_fwd', mutates_args=())
def layer_fwd(u_prev: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], w_fg: Tensor, b_fg: Tensor,
              w_mlp: Optional[Tensor], b_mlp: Optional[Tensor], supports: List[Tensor],
              drop_mask: Optional[Tensor], rng: Optional[Tensor], hop_mats: Optional[Tensor], Lf: int, taps: int,
              dilation: int, order: int, training: bool, has_gconv: bool, dropout_p: float, seed: int, offset: int
              ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    return _layer_fwd_impl(u_prev, scale, shift, w_fg, b_fg, w_mlp, b_mlp, supports, drop_mask, rng, hop_mats, Lf, taps,
                           dilation, order, training, has_gconv, dropout_p, seed, offset)


@torch.library.custom_op('gwn::layer_fwd_bn', mutates_args=('rmean', 'rvar'))
def layer_fwd_bn(u_prev: Tensor, bn_stats: Optional[Tensor], count: float, gamma: Tensor, beta: Tensor, rmean: Tensor,
                 rvar: Tensor, momentum: float, eps: float, w_fg: Tensor, b_fg: Tensor,
                 w_mlp: Optional[Tensor], b_mlp: Optional[Tensor], supports: List[Tensor],
                 drop_mask: Optional[Tensor], rng: Optional[Tensor], hop_mats: Tensor, Lf: int, taps: int,
                 dilation: int, order: int, training: bool, has_gconv: bool, dropout_p: float, seed: int, offset: int
                 ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """layer_fwd with the previous layer's BatchNorm (graph_wavenet.py:250) folded inside the gate kernel's prologue
    (bf16 tensor-core path): batch statistics `bn_stats` [2,32] (training) or the running statistics (eval) ->
    scale / shift / mean / rstd come back as outputs, the running statistics are updated in place."""
    f32 = dict(device=u_prev.device, dtype=torch.float32)
    scale, shift, mean, rstd = (torch.empty((CH,), **f32) for _ in range(4))
    if training:
        _req(bn_stats, torch.float64, 'bn_stats')
    out = _layer_fwd_impl(u_prev, scale, shift, w_fg, b_fg, w_mlp, b_mlp, supports, drop_mask, rng, hop_mats, Lf, taps,
                          dilation, order, training, has_gconv, dropout_p, seed, offset,
                          bn=(bn_stats if training else None, count, gamma, beta, rmean, rvar, momentum, eps, mean, rstd))
    return (*out, scale, shift, mean, rstd)


@layer_fwd_bn.register_fake
def _(u_prev, bn_stats, count, gamma, beta, rmean, rvar, momentum, eps, w_fg, b_fg, w_mlp, b_mlp, supports, drop_mask,
      rng, hop_mats, Lf, taps, dilation, order, training, has_gconv, dropout_p, seed, offset):
    N, Lin, V, _c = u_prev.shape
    Lout = Lin - dilation * (taps - 1)
    full = (N, Lout, V, CH)
    v = lambda: u_prev.new_empty((CH,), dtype=torch.float32)  # noqa: E731
    return (u_prev.new_empty(full if has_gconv else (0,)), u_prev.new_empty((2, CH), dtype=torch.float64),
            u_prev.new_empty((N, Lf, V, CH)), u_prev.new_empty(full if training else (0,)),
            u_prev.new_empty(full if training else (0,)), v(), v(), v(), v())


@layer_fwd.register_fake
def _(u_prev, scale, shift, w_fg, b_fg, w_mlp, b_mlp, supports, drop_mask, rng, hop_mats, Lf, taps, dilation, order,
      training, has_gconv, dropout_p, seed, offset):
    N, Lin, V, _c = u_prev.shape
    Lout = Lin - dilation * (taps - 1)
    full = (N, Lout, V, CH)
    return (u_prev.new_empty(full if has_gconv else (0,)), u_prev.new_empty((2, CH), dtype=torch.float64),
            u_prev.new_empty((N, Lf, V, CH)), u_prev.new_empty(full if training else (0,)),
            u_prev.new_empty(full if training else (0,)))


def _layer_bwd_impl(u_prev, scale, shift, w_fg, w_mlp, supports, needs_grad, drop_mask, rng, hop_mats, a, b, du, dz_last,
                    Lf, taps, dilation, order, training, dropout_p, seed, offset, flat_ws=None, d_acc=None):
    """flat_ws / d_acc: optional caller-owned, already zeroed accumulation buffers (the model hands every layer a slice
    of ONE per-step buffer and ONE shared support-gradient accumulator: no per-layer fill, no per-layer add)."""
    N, Lin, V, _c = u_prev.shape
    Lout = Lin - dilation * (taps - 1)
    dt, dev = u_prev.dtype, u_prev.device
    has_du = du is not None
    n_sup = len(supports)
    mlp_in = CH * (1 + order * n_sup)
    P = N * Lout * V
    f32 = dict(device=dev, dtype=torch.float32)
    # bf16 tensor-core path: dx_prev only lives until the BatchNorm backward (or the start conv backward) reads it, so it is
    # stored in bf16 (its BN statistics come from the fp32 accumulator inside the kernel)
    dx_bf16 = dt == torch.bfloat16 and hop_mats is not None and taps <= 4
    dx_prev = torch.empty((N, Lin, V, CH), device=dev, dtype=torch.bfloat16 if dx_bf16 else torch.float32)
    # the small accumulated outputs live in ONE zero-filled buffer (one fill kernel instead of six memset nodes)
    pairs = [s.dim() == 3 for s in supports]         # [2,V,V] supports take their gradient in factored form (d0, Q)
    if flat_ws is None:
        sizes = _layer_bwd_sizes(taps, mlp_in, V, needs_grad, has_du, pairs)
        flat = torch.zeros((sum(sizes),), **f32)
    else:                                            # the support gradient goes to the shared accumulator
        sizes = _layer_bwd_sizes(taps, mlp_in, V, [False] * len(needs_grad), has_du, pairs)
        flat = flat_ws
        if flat.numel() < sum(sizes) or flat.dtype != torch.float32:
            raise ValueError('layer_bwd: gradient workspace too small')
    parts, off = [], 0
    for sz in sizes:
        parts.append(flat[off:off + sz]); off += sz
    dx_stats = parts[0].view(torch.float64).view(2, CH)
    dw_fg, db_fg = parts[1].view(taps * CH, 2 * CH), parts[2]
    dw_mlp, db_mlp = parts[3].view(mlp_in, CH), parts[4]
    if flat_ws is None:
        d_sup = [parts[5 + i] if (g and has_du) else torch.empty((0,), **f32) for i, g in enumerate(needs_grad)]
    else:
        d_sup = [d_acc.reshape(-1) if (g and has_du) else torch.empty((0,), **f32) for g in needs_grad]
    ws_cat = torch.empty((P, mlp_in) if has_du else (0,), device=dev, dtype=dt)
    ws_dcat = torch.empty((P, mlp_in) if has_du else (0,), device=dev, dtype=dt)
    ws_dfg = torch.empty((P, 2 * CH), **f32)
    ws_w = torch.empty((128 * 1024,), device=dev, dtype=torch.uint8) if hop_mats is not None else None
    cfg = _layer_cfg(N, V, Lin, Lout, Lf, taps, dilation, n_sup, order, dt, training, True, dropout_p, seed, offset)
    args = LayerBwdArgs(u_prev=_p(u_prev), scale=_p(scale), shift=_p(shift), w_fg=_p(w_fg), w_mlp=_p(w_mlp),
                        supports=_sup_array(supports), drop_mask=_p(drop_mask), rng=_p(rng),
                        hop_mats=_p(hop_mats), ws_w=_p(ws_w), a=_p(a), b=_p(b),
                        du=_p(du), dz_last=_p(dz_last), dx_prev=_p(dx_prev), dx_stats=_p(dx_stats),
                        dw_fg=_p(dw_fg), db_fg=_p(db_fg), dw_mlp=_p(dw_mlp), db_mlp=_p(db_mlp),
                        ws_cat=_p(ws_cat) if has_du else None, ws_dcat=_p(ws_dcat) if has_du else None,
                        ws_dfg=_p(ws_dfg), outputs_zeroed=1, dx_prev_bf16=int(dx_bf16))
    ell, _ell_keep = _ell_array(supports) if (has_du and hop_mats is not None and dt == torch.bfloat16) else (None, None)
    if ell is not None:
        args.ell = ell
    for i, g in enumerate(needs_grad):
        args.support_needs_grad[i] = int(bool(g) and has_du)
        args.d_supports[i] = d_sup[i].data_ptr() if (g and has_du) else None
        args.d_supports_sq[i] = d_sup[i][V * V:].data_ptr() if (g and has_du and pairs[i]) else None
    with torch.cuda.device(dev):
        check(lib().gwn_layer_bwd(C.byref(cfg), C.byref(args), _stream()), 'gwn_layer_bwd')
    return [dx_prev, flat]


@torch.library.custom_op('gwn::layer_bwd', mutates_args=())
def layer_bwd(u_prev: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], w_fg: Tensor,
              w_mlp: Optional[Tensor], supports: List[Tensor], needs_grad: List[bool],
              drop_mask: Optional[Tensor], rng: Optional[Tensor], hop_mats: Optional[Tensor], a: Tensor, b: Tensor,
              du: Optional[Tensor], dz_last: Optional[Tensor], Lf: int, taps: int, dilation: int, order: int, training: bool,
              dropout_p: float, seed: int, offset: int) -> List[Tensor]:
    """Returns [dx_prev, flat]: flat is ONE zero-initialised fp32 buffer holding, back to back, dx_stats (64 fp64),
    dw_fg, db_fg, dw_mlp, db_mlp and d_support_i for the supports whose needs_grad is True - see
    `_split_layer_bwd` (custom-op outputs may not alias, so the views are cut outside the op)."""
    return _layer_bwd_impl(u_prev, scale, shift, w_fg, w_mlp, supports, needs_grad, drop_mask, rng, hop_mats, a, b, du,
                           dz_last, Lf, taps, dilation, order, training, dropout_p, seed, offset)


@torch.library.custom_op('gwn::layer_bwd_ws', mutates_args=('flat_ws', 'd_acc'))
def layer_bwd_ws(flat_ws: Tensor, d_acc: Optional[Tensor], u_prev: Tensor, scale: Optional[Tensor], shift: Optional[Tensor],
                 w_fg: Tensor, w_mlp: Optional[Tensor], supports: List[Tensor], needs_grad: List[bool],
                 drop_mask: Optional[Tensor], rng: Optional[Tensor], hop_mats: Optional[Tensor], a: Tensor, b: Tensor,
                 du: Optional[Tensor], dz_last: Optional[Tensor], Lf: int, taps: int, dilation: int, order: int,
                 training: bool, dropout_p: float, seed: int, offset: int) -> Tensor:
    """layer_bwd accumulating into caller-owned zeroed buffers: `flat_ws` (this layer's slice of the per-step gradient
    workspace: dx_stats, dw_fg, db_fg, dw_mlp, db_mlp) and `d_acc` (the ONE support-gradient accumulator every layer of
    the step adds into; at most one support may need a gradient).  Returns dx_prev."""
    if sum(bool(g) for g in needs_grad) > (1 if d_acc is not None else 0):
        raise ValueError('layer_bwd_ws: one shared accumulator serves one support gradient')
    return _layer_bwd_impl(u_prev, scale, shift, w_fg, w_mlp, supports, needs_grad, drop_mask, rng, hop_mats, a, b, du,
                           dz_last, Lf, taps, dilation, order, training, dropout_p, seed, offset, flat_ws=flat_ws,
                           d_acc=d_acc)[0]


@layer_bwd_ws.register_fake
def _(flat_ws, d_acc, u_prev, scale, shift, w_fg, w_mlp, supports, needs_grad, drop_mask, rng, hop_mats, a, b, du, dz_last,
      Lf, taps, dilation, order, training, dropout_p, seed, offset):
    N, Lin, V, _c = u_prev.shape
    dx_bf16 = u_prev.dtype == torch.bfloat16 and hop_mats is not None and taps <= 4
    return u_prev.new_empty((N, Lin, V, CH), dtype=torch.bfloat16 if dx_bf16 else torch.float32)


def _layer_bwd_sizes(taps: int, mlp_in: int, V: int, needs_grad: Sequence[bool], has_du: bool,
                     pairs: Optional[Sequence[bool]] = None) -> List[int]:
    pairs = pairs if pairs is not None else [False] * len(needs_grad)
    return [2 * 2 * CH, taps * CH * 2 * CH, 2 * CH, mlp_in * CH, CH] + \
        [(2 if pr else 1) * V * V if (g and has_du) else 0 for g, pr in zip(needs_grad, pairs)]


def _split_layer_bwd(flat: Tensor, taps: int, mlp_in: int, V: int, needs_grad: Sequence[bool], has_du: bool,
                     pairs: Optional[Sequence[bool]] = None):
    pairs = pairs if pairs is not None else [False] * len(needs_grad)
    sizes = _layer_bwd_sizes(taps, mlp_in, V, needs_grad, has_du, pairs)
    parts, off = [], 0
    for sz in sizes:
        parts.append(flat[off:off + sz]); off += sz
    dx_stats = parts[0].view(torch.float64).view(2, CH)
    d_sup = [(parts[5 + i].view(2, V, V) if pairs[i] else parts[5 + i].view(V, V)) if sizes[5 + i] else None
             for i in range(len(needs_grad))]
    return dx_stats, parts[1].view(taps * CH, 2 * CH), parts[2], parts[3].view(mlp_in, CH), parts[4], d_sup


@layer_bwd.register_fake
def _(u_prev, scale, shift, w_fg, w_mlp, supports, needs_grad, drop_mask, rng, hop_mats, a, b, du, dz_last, Lf,
      taps, dilation, order, training, dropout_p, seed, offset):
    N, Lin, V, _c = u_prev.shape
    mlp_in = CH * (1 + order * len(supports))
    f = lambda *s: u_prev.new_empty(s, dtype=torch.float32)  # noqa: E731
    dx_bf16 = u_prev.dtype == torch.bfloat16 and hop_mats is not None and taps <= 4
    return [u_prev.new_empty((N, Lin, V, CH), dtype=torch.bfloat16 if dx_bf16 else torch.float32),
            f(sum(_layer_bwd_sizes(taps, mlp_in, V, needs_grad, du is not None, [s.dim() == 3 for s in supports])))]


class WaveNetLayer(torch.autograd.Function):
    """bn_prev-fold -> gated dilated conv -> (z_last) -> diffusion conv + dropout + residual -> (u, BN stats).

    Inputs : u_prev, stats_prev|None, gamma_prev|None, beta_prev|None, rmean_prev|None, rvar_prev|None,
             w_fg, b_fg, w_mlp|None, b_mlp|None, drop_mask|None, rng|None, hop_mats|None, meta(dict), *supports
    Outputs: u (pre-BN, or empty if the gconv is skipped), stats (non-differentiable), z_last
    """

    @staticmethod
    def forward(ctx, u_prev, stats_prev, gamma, beta, rmean, rvar, w_fg, b_fg, w_mlp, b_mlp, drop_mask, rng, hop_mats,
                meta, *supports):
        m = meta
        training = m['training']
        has_bn = gamma is not None
        scale = shift = mean = rstd = None
        count = float(u_prev.numel() // CH)
        sup = [s.contiguous() for s in supports]
        # bf16 tensor-core path: the BatchNorm fold happens inside the gate kernel's prologue (no fold / prep launch)
        fused_bn = has_bn and u_prev.dtype == torch.bfloat16 and hop_mats is not None and m['taps'] <= 4
        if fused_bn:
            u, stats, z_last, a, b, scale, shift, mean, rstd = layer_fwd_bn(
                u_prev, stats_prev if training else None, count, gamma, beta, rmean, rvar, m['momentum'], m['eps'],
                w_fg, b_fg, w_mlp, b_mlp, sup, drop_mask, rng, hop_mats, m['Lf'], m['taps'], m['dilation'], m['order'],
                training, m['has_gconv'], m['dropout_p'], m['seed'], m['offset'])
        else:
            if has_bn:
                scale, shift, mean, rstd = bn_fold(stats_prev if training else None, count, gamma, beta, rmean, rvar,
                                                   m['momentum'], m['eps'], training)
            u, stats, z_last, a, b = layer_fwd(u_prev, scale, shift, w_fg, b_fg, w_mlp, b_mlp, sup, drop_mask, rng,
                                               hop_mats, m['Lf'], m['taps'], m['dilation'], m['order'], training,
                                               m['has_gconv'], m['dropout_p'], m['seed'], m['offset'])
        ctx.set_materialize_grads(False)      # a dead output (last layer's u) must stay "no gradient"
        ctx.meta = m
        ctx.count = count
        ctx.has_bn = has_bn
        ctx.n_sup = len(sup)
        ctx.sup_needs = [bool(s.requires_grad) for s in supports]
        ctx.save_for_backward(u_prev, scale, shift, mean, rstd, gamma, w_fg, w_mlp, drop_mask, rng, hop_mats, a, b,
                              *sup)
        ctx.mark_non_differentiable(stats)
        return u, stats, z_last

    @staticmethod
    def backward(ctx, du, _dstats, dz_last):
        m = ctx.meta
        (u_prev, scale, shift, mean, rstd, gamma, w_fg, w_mlp, drop_mask, rng, hop_mats, a, b,
         *sup) = ctx.saved_tensors
        if not m['training']:
            raise _lib.GwnError('backward through an eval-mode forward is not supported (a,b were not saved)')
        if du is not None and du.numel() == 0:
            du = None
        has_du = du is not None and m['has_gconv']
        gw = m.get('grad_ws')
        n_sup_b = len(sup) if has_du else 0
        needs = ctx.sup_needs if has_du else []
        args = (u_prev, scale, shift, w_fg, w_mlp if has_du else None, sup if has_du else [], needs, drop_mask, rng,
                hop_mats, a, b, du.contiguous() if has_du else None,
                dz_last.contiguous() if dz_last is not None else None,
                m['Lf'], m['taps'], m['dilation'], m['order'], True, m['dropout_p'], m['seed'], m['offset'])
        if gw is not None and sum(needs) <= 1:
            # per-step gradient workspace (graph_wavenet._forward_nchw): this layer's zeroed slice + the shared
            # support-gradient accumulator; the accumulated support gradient is handed to autograd by layer 0 only
            # (its backward runs last: every other layer's contribution is already in)
            flat = gw['flat']
            dx_prev = layer_bwd_ws(flat, gw['d_acc'] if any(needs) else None, *args)
            dx_stats, dw_fg, db_fg, dw_mlp, db_mlp, _none = _split_layer_bwd(
                flat, m['taps'], CH * (1 + m['order'] * n_sup_b), u_prev.shape[2], [False] * len(needs), has_du,
                [s.dim() == 3 for s in sup] if has_du else [])
            d_sup = [gw['d_acc'] if (g and gw['first']) else None for g in needs]
        else:
            dx_prev, flat = layer_bwd(*args)
            dx_stats, dw_fg, db_fg, dw_mlp, db_mlp, d_sup = _split_layer_bwd(
                flat, m['taps'], CH * (1 + m['order'] * n_sup_b), u_prev.shape[2], needs, has_du,
                [s.dim() == 3 for s in sup] if has_du else [])
        if ctx.has_bn:
            du_prev, dgamma, dbeta = bn_bwd(dx_prev, u_prev, dx_stats, ctx.count, gamma, mean, rstd, True)
        else:
            du_prev, dgamma, dbeta = dx_prev.to(u_prev.dtype), None, None
        g_sup = []
        for i in range(ctx.n_sup):
            g_sup.append(d_sup[i] if (has_du and ctx.sup_needs[i]) else None)
        return (du_prev, None, dgamma, dbeta, None, None, dw_fg, db_fg,
                dw_mlp if has_du else None, db_mlp if has_du else None, None, None, None, None, *g_sup)


# =========================================================================== head (:231-236, :252-254)
def _ptr_array(ts: Sequence[Tensor]):
    arr = (C.c_void_p * _lib.MAX_LAYERS)()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return arr


@torch.library.custom_op('gwn::head_fwd', mutates_args=())
def head_fwd(z_last: List[Tensor], w_skip: Tensor, b_skip: Tensor, w_end1: Tensor, b_end1: Tensor, w_end2: Tensor,
             b_end2: Tensor, out_dim: int) -> Tuple[Tensor, Tensor, Tensor]:
    z0 = _req(z_last[0], None, 'z_last')
    N, Lf, V, _c = z0.shape
    S, E, Opad = w_skip.shape[1], w_end1.shape[1], w_end2.shape[1]
    P = N * Lf * V
    f32 = dict(device=z0.device, dtype=torch.float32)
    s1, e1 = torch.empty((P, S), **f32), torch.empty((P, E), **f32)
    out = torch.empty((N, out_dim, V, Lf), **f32)
    ws = torch.empty((P, Opad), **f32)
    cfg = HeadCfg(N=N, V=V, Lf=Lf, n_layers=len(z_last), S=S, E=E, O=out_dim, dtype=_code(z0.dtype))
    args = HeadFwdArgs(z_last=_ptr_array(z_last), w_skip=_p(w_skip), b_skip=_p(b_skip), w_end1=_p(w_end1),
                       b_end1=_p(b_end1), w_end2=_p(w_end2), b_end2=_p(b_end2), s1=_p(s1), e1=_p(e1),
                       out=_p(out), ws=_p(ws))
    with torch.cuda.device(z0.device):
        check(lib().gwn_head_fwd(C.byref(cfg), C.byref(args), _stream()), 'gwn_head_fwd')
    return out, s1, e1


@head_fwd.register_fake
def _(z_last, w_skip, b_skip, w_end1, b_end1, w_end2, b_end2, out_dim):
    N, Lf, V, _c = z_last[0].shape
    f = lambda *s: z_last[0].new_empty(s, dtype=torch.float32)  # noqa: E731
    return f(N, out_dim, V, Lf), f(N * Lf * V, w_skip.shape[1]), f(N * Lf * V, w_end1.shape[1])


@torch.library.custom_op('gwn::head_bwd', mutates_args=())
def head_bwd(z_last: List[Tensor], w_skip: Tensor, w_end1: Tensor, w_end2: Tensor, s1: Tensor, e1: Tensor,
             dout: Tensor) -> List[Tensor]:
    """Returns [dw_skip, db_skip, dw_end1, db_end1, dw_end2, db_end2, dz_last_0, ...]."""
    z0 = z_last[0]
    N, Lf, V, _c = z0.shape
    S, E, Opad = w_skip.shape[1], w_end1.shape[1], w_end2.shape[1]
    O = dout.shape[1]
    P = N * Lf * V
    f32 = dict(device=z0.device, dtype=torch.float32)
    dw_skip, db_skip = torch.empty_like(w_skip), torch.empty((S,), **f32)
    dw_end1, db_end1 = torch.empty_like(w_end1), torch.empty((E,), **f32)
    dw_end2, db_end2 = torch.empty_like(w_end2), torch.empty((Opad,), **f32)
    dz = [torch.empty_like(z) for z in z_last]
    ws_do, ws_de1, ws_ds1 = torch.empty((P, Opad), **f32), torch.empty((P, E), **f32), torch.empty((P, S), **f32)
    cfg = HeadCfg(N=N, V=V, Lf=Lf, n_layers=len(z_last), S=S, E=E, O=O, dtype=_code(z0.dtype))
    args = HeadBwdArgs(z_last=_ptr_array(z_last), w_skip=_p(w_skip), w_end1=_p(w_end1), w_end2=_p(w_end2),
                       s1=_p(s1), e1=_p(e1), dout=_p(dout), dw_skip=_p(dw_skip), db_skip=_p(db_skip),
                       dw_end1=_p(dw_end1), db_end1=_p(db_end1), dw_end2=_p(dw_end2), db_end2=_p(db_end2),
                       dz_last=_ptr_array(dz), ws_do=_p(ws_do), ws_de1=_p(ws_de1), ws_ds1=_p(ws_ds1))
    with torch.cuda.device(z0.device):
        check(lib().gwn_head_bwd(C.byref(cfg), C.byref(args), _stream()), 'gwn_head_bwd')
    return [dw_skip, db_skip, dw_end1, db_end1, dw_end2, db_end2] + dz


@head_bwd.register_fake
def _(z_last, w_skip, w_end1, w_end2, s1, e1, dout):
    f = lambda t: torch.empty_like(t)  # noqa: E731
    z0 = z_last[0]
    v = lambda n: z0.new_empty((n,), dtype=torch.float32)  # noqa: E731
    return [f(w_skip), v(w_skip.shape[1]), f(w_end1), v(w_end1.shape[1]), f(w_end2), v(w_end2.shape[1])] + \
        [f(z) for z in z_last]


@torch.library.custom_op('gwn::head_fwd_tc', mutates_args=())
def head_fwd_tc(zcat: Tensor, w_skip: Tensor, b_skip: Tensor, w_end1: Tensor, b_end1: Tensor, w_end2: Tensor,
                b_end2: Tensor, out_dim: int) -> Tuple[Tensor, Tensor, Tensor]:
    """bf16 head on the tensor cores (csrc/head_tc.cu).  zcat: [N,Lf,V,32*n_layers] bf16."""
    _req(zcat, torch.bfloat16, 'zcat')
    N, Lf, V, K0 = zcat.shape
    S, E = w_skip.shape[1], w_end1.shape[1]
    P = N * Lf * V
    dev = zcat.device
    s1 = torch.empty((P, 2 * S), device=dev, dtype=torch.bfloat16)      # [hi | lo] split (csrc/head_tc.cu)
    e1 = torch.empty((P, E), device=dev, dtype=torch.bfloat16)
    out = torch.empty((N, out_dim, V, Lf), device=dev, dtype=torch.float32)
    ws_w = torch.empty((lib().gwn_head_tc_ws_bytes(K0 // CH, S, E, out_dim),), device=dev, dtype=torch.uint8)
    cfg = HeadCfg(N=N, V=V, Lf=Lf, n_layers=K0 // CH, S=S, E=E, O=out_dim, dtype=GWN_BF16)
    args = HeadTcFwdArgs(zcat=_p(zcat), w_skip=_p(w_skip), b_skip=_p(b_skip), w_end1=_p(w_end1), b_end1=_p(b_end1),
                         w_end2=_p(w_end2), b_end2=_p(b_end2), s1=_p(s1), e1=_p(e1), out=_p(out), ws_w=_p(ws_w))
    with torch.cuda.device(dev):
        check(lib().gwn_head_fwd_tc(C.byref(cfg), C.byref(args), _stream()), 'gwn_head_fwd_tc')
    return out, s1, e1


@head_fwd_tc.register_fake
def _(zcat, w_skip, b_skip, w_end1, b_end1, w_end2, b_end2, out_dim):
    N, Lf, V, _k = zcat.shape
    P = N * Lf * V
    return (zcat.new_empty((N, out_dim, V, Lf), dtype=torch.float32), zcat.new_empty((P, 2 * w_skip.shape[1])),
            zcat.new_empty((P, w_end1.shape[1])))


@torch.library.custom_op('gwn::head_bwd_tc', mutates_args=())
def head_bwd_tc(zcat: Tensor, w_skip: Tensor, w_end1: Tensor, w_end2: Tensor, s1: Tensor, e1: Tensor,
                dout: Tensor) -> List[Tensor]:
    """Returns [flat, dz_last_0, ...]: flat = dw_skip | db_skip | dw_end1 | db_end1 | dw_end2 | db_end2 back to back
    (`_split_head_bwd`); dz_last_i: [N,Lf,V,32] bf16."""
    N, Lf, V, K0 = zcat.shape
    nl = K0 // CH
    S, E, Opad = w_skip.shape[1], w_end1.shape[1], w_end2.shape[1]
    O = dout.shape[1]
    P = N * Lf * V
    dev = zcat.device
    f32 = dict(device=dev, dtype=torch.float32)
    b16 = dict(device=dev, dtype=torch.bfloat16)
    # the six parameter gradients share ONE zero-filled buffer (one fill instead of six memset nodes); cut into views
    # by `_split_head_bwd` outside the op (custom-op outputs may not alias)
    sizes = _head_bwd_sizes(K0, S, E, Opad)
    flat = torch.zeros((sum(sizes),), **f32)
    offs = [sum(sizes[:i]) for i in range(6)]
    dw_skip, db_skip, dw_end1, db_end1, dw_end2, db_end2 = (flat[o:o + sz] for o, sz in zip(offs, sizes))
    dz = [torch.empty((N, Lf, V, CH), **b16) for _ in range(nl)]
    ws_do, ws_de1, ws_ds1 = torch.empty((P, Opad), **b16), torch.empty((P, E), **b16), torch.empty((P, S), **b16)
    ws_w = torch.empty((lib().gwn_head_tc_ws_bytes(nl, S, E, O),), device=dev, dtype=torch.uint8)
    cfg = HeadCfg(N=N, V=V, Lf=Lf, n_layers=nl, S=S, E=E, O=O, dtype=GWN_BF16)
    args = HeadTcBwdArgs(zcat=_p(zcat), w_skip=_p(w_skip), w_end1=_p(w_end1), w_end2=_p(w_end2), s1=_p(s1), e1=_p(e1),
                         dout=_p(dout), dw_skip=_p(dw_skip), db_skip=_p(db_skip), dw_end1=_p(dw_end1),
                         db_end1=_p(db_end1), dw_end2=_p(dw_end2), db_end2=_p(db_end2), dz_last=_ptr_array(dz),
                         ws_do=_p(ws_do), ws_de1=_p(ws_de1), ws_ds1=_p(ws_ds1), ws_w=_p(ws_w), outputs_zeroed=1)
    with torch.cuda.device(dev):
        check(lib().gwn_head_bwd_tc(C.byref(cfg), C.byref(args), _stream()), 'gwn_head_bwd_tc')
    return [flat] + dz


def _head_bwd_sizes(K0: int, S: int, E: int, Opad: int) -> List[int]:
    return [K0 * S, S, S * E, E, E * Opad, Opad]


def _split_head_bwd(flat: Tensor, K0: int, S: int, E: int, Opad: int):
    sizes = _head_bwd_sizes(K0, S, E, Opad)
    offs = [sum(sizes[:i]) for i in range(6)]
    p = [flat[o:o + sz] for o, sz in zip(offs, sizes)]
    return [p[0].view(K0, S), p[1], p[2].view(S, E), p[3], p[4].view(E, Opad), p[5]]


@head_bwd_tc.register_fake
def _(zcat, w_skip, w_end1, w_end2, s1, e1, dout):
    N, Lf, V, K0 = zcat.shape
    S, E, Opad = w_skip.shape[1], w_end1.shape[1], w_end2.shape[1]
    return [zcat.new_empty((sum(_head_bwd_sizes(K0, S, E, Opad)),), dtype=torch.float32)] + \
        [zcat.new_empty((N, Lf, V, CH)) for _ in range(K0 // CH)]


def head_tc_supported(P: int, S: int, E: int) -> bool:
    """TMA-fed tensor-core head: needs channel counts TMA can box (multiples of 64 keep every box full)."""
    return S % 64 == 0 and E % 64 == 0 and P < 2 ** 31


# parity-test hook: a dict here receives the saved head activations of the next bf16 forward ('s1': [P, hi | lo] of
# relu(skip), 'e1': relu(end_conv_1)) - the ReLU decisions the backward will use (tests/gpu_helpers.captured_head_masks)
HEAD_CAPTURE: Optional[dict] = None


class SkipHead(torch.autograd.Function):
    """relu(sum_i Ws_i z_i[..., -Lf:] + sum_i bs_i) -> relu(end_conv_1) -> end_conv_2, NCHW out."""

    @staticmethod
    def forward(ctx, w_skip, b_skip, w_end1, b_end1, w_end2, b_end2, out_dim, *z_last):
        z0 = z_last[0]
        ctx.tc = (z0.dtype == torch.bfloat16 and
                  head_tc_supported(z0.numel() // CH, w_skip.shape[1], w_end1.shape[1]))
        if ctx.tc:
            zcat = torch.cat(z_last, dim=-1)
            out, s1, e1 = head_fwd_tc(zcat, w_skip, b_skip, w_end1, b_end1, w_end2, b_end2, out_dim)
            ctx.save_for_backward(w_skip, w_end1, w_end2, s1, e1, zcat)
            if HEAD_CAPTURE is not None:
                HEAD_CAPTURE.update(s1=s1.detach().clone(), e1=e1.detach().clone())
            return out
        zs = [z.contiguous() for z in z_last]
        out, s1, e1 = head_fwd(zs, w_skip, b_skip, w_end1, b_end1, w_end2, b_end2, out_dim)
        ctx.save_for_backward(w_skip, w_end1, w_end2, s1, e1, *zs)
        return out

    @staticmethod
    def backward(ctx, dout):
        w_skip, w_end1, w_end2, s1, e1, *zs = ctx.saved_tensors
        if ctx.tc:
            outs = head_bwd_tc(zs[0], w_skip, w_end1, w_end2, s1, e1, dout.contiguous().float())
            outs = _split_head_bwd(outs[0], w_skip.shape[0], w_skip.shape[1], w_end1.shape[1], w_end2.shape[1]) + \
                list(outs[1:])
        else:
            outs = head_bwd(list(zs), w_skip, w_end1, w_end2, s1, e1, dout.contiguous().float())
        return (*outs[:6], None, *outs[6:])


# =========================================================================== training loss (lit.py:24)
@torch.library.custom_op('gwn::mse_loss_fwd', mutates_args=())
def mse_loss_fwd(a: Tensor, b: Tensor) -> Tensor:
    """mean((a - b)^2) of two fp32 CUDA tensors of the same shape as a 0-dim tensor, one launch."""
    _req(a, torch.float32, 'input'); _req(b, torch.float32, 'target')
    loss = torch.empty((), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        check(lib().gwn_mse_loss_fwd(_p(a), _p(b), a.numel(), _p(loss), _stream()), 'gwn_mse_loss_fwd')
    return loss


@mse_loss_fwd.register_fake
def _(a, b):
    return a.new_empty(())


@torch.library.custom_op('gwn::mse_loss_bwd', mutates_args=())
def mse_loss_bwd(a: Tensor, b: Tensor, grad_loss: Tensor) -> Tensor:
    _req(a, torch.float32, 'input'); _req(b, torch.float32, 'target'); _req(grad_loss, torch.float32, 'grad_loss')
    da = torch.empty_like(a)
    with torch.cuda.device(a.device):
        check(lib().gwn_mse_loss_bwd(_p(a), _p(b), _p(grad_loss), a.numel(), _p(da), _stream()), 'gwn_mse_loss_bwd')
    return da


@mse_loss_bwd.register_fake
def _(a, b, grad_loss):
    return torch.empty_like(a)


class _MSELoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return mse_loss_fwd(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return mse_loss_bwd(a, b, g.contiguous()), None


def mse_loss(input: Tensor, target: Tensor) -> Tensor:
    """``nn.MSELoss()(input, target)`` (lit.py:24; reduction 'mean', gradient for `input` only) in one forward and one
    backward launch for fp32 CUDA tensors; anything else (other dtypes, a target that needs a gradient, shapes that would
    broadcast) goes to ``torch.nn.functional.mse_loss`` unchanged."""
    if (input.is_cuda and target.is_cuda and input.dtype == torch.float32 and target.dtype == torch.float32 and
            input.shape == target.shape and not target.requires_grad and 0 < input.numel() < 2 ** 31):
        return _MSELoss.apply(input.contiguous(), target.contiguous())
    return torch.nn.functional.mse_loss(input, target)


# =========================================================================== nconv primitive (:60-66)
@torch.library.custom_op('gwn::node_mix', mutates_args=())
def node_mix(x: Tensor, A: Tensor, transpose_a: bool) -> Tensor:
    """x: [slabs, V, 32] channels-last; returns y[s,w,c] = sum_v x[s,v,c] * A[v,w] (or A[w,v])."""
    _req(x, None, 'x'); _req(A, torch.float32, 'A')
    S, V, _c = x.shape
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib().gwn_node_mix(_p(x), CH, 0, _p(y), CH, 0, 0, _p(A), int(transpose_a), S, V, _code(x.dtype),
                                 _stream()), 'gwn_node_mix')
    return y


@node_mix.register_fake
def _(x, A, transpose_a):
    return torch.empty_like(x)


# =========================================================================== tcgen05 diffusion hops (V <= 128, bf16)
def hop_mode(V: int, n_supports: int) -> int:
    """Image format the layer kernels expect in ``hop_mats`` (include/gwn.h: gwn_hop_mode - the C side applies the same
    function): 1 = `hop_mats` images (supports resident in shared memory), 2 = `support_images` (TMA-tiled hop GEMMs)."""
    m = lib().gwn_hop_mode(int(V), int(n_supports))
    if m not in (1, 2):
        raise ValueError(f'no tensor-core hop path for V={V}, {n_supports} supports')
    return m


def hop_tc_supported(V: int, n_supports: int = 3) -> bool:
    """True when every support image stays resident in shared memory (mode 1)."""
    return hop_mode(V, n_supports) == 1


@torch.library.custom_op('gwn::hop_mats', mutates_args=())
def hop_mats(supports: List[Tensor]) -> Tensor:
    """UMMA A-operand images of every support (A^T, (A^2)^T, A, A^2 per support), built once per forward."""
    for s in supports:
        _req(s, torch.float32, 'support')
    V = supports[0].shape[0]
    nbytes = lib().gwn_hop_mats_bytes(V, len(supports))
    out = torch.empty((nbytes // 2,), device=supports[0].device, dtype=torch.bfloat16)
    ptrs = (C.c_void_p * len(supports))(*[s.data_ptr() for s in supports])
    with torch.cuda.device(out.device):
        check(lib().gwn_hop_mats_prep(ptrs, len(supports), V, _p(out), _stream()), 'gwn_hop_mats_prep')
    return out


@hop_mats.register_fake
def _(supports):
    V = supports[0].shape[0]
    return supports[0].new_empty((lib().gwn_hop_mats_bytes(V, len(supports)) // 2,), dtype=torch.bfloat16)


@torch.library.custom_op('gwn::support_images', mutates_args=())
def support_images(supports: List[Tensor]) -> Tensor:
    """bf16 operand images of every support for the TMA-tiled hop GEMM (graphs too big to keep on chip):
    [n_supports][2][V][Vp], image 0 = A^T (forward hop), image 1 = A (backward hop)."""
    for s in supports:
        _req(s, torch.float32, 'support')
    V = supports[0].shape[0]
    nbytes = lib().gwn_support_images_bytes(V, len(supports))
    out = torch.empty((nbytes // 2,), device=supports[0].device, dtype=torch.bfloat16)
    ptrs = (C.c_void_p * len(supports))(*[s.data_ptr() for s in supports])
    with torch.cuda.device(out.device):
        check(lib().gwn_support_images_prep(ptrs, len(supports), V, _p(out), _stream()), 'gwn_support_images_prep')
    return out


@support_images.register_fake
def _(supports):
    V = supports[0].shape[0]
    Vp = 8 * ((V + 7) // 8)
    return supports[0].new_empty((len(supports) * 2 * V * Vp,), dtype=torch.bfloat16)


@torch.library.custom_op('gwn::hop_tc', mutates_args=('buf',))
def hop_tc(mats: Tensor, n_mats: int, mat: int, buf: Tensor, slot_in: int, slot_out: int, V: int) -> None:
    """One tensor-core hop inside a pitched bf16 buffer [slabs*V, pitch]: slot_out <- Mop[mat] * slot_in."""
    _req(buf, torch.bfloat16, 'buf')
    rows, pitch = buf.shape
    with torch.cuda.device(buf.device):
        check(lib().gwn_hop_tc(_p(mats), n_mats, mat, _p(buf), pitch, slot_in, slot_out, rows // V, V, _stream()),
              'gwn_hop_tc')


# =========================================================================== parameter packing
def _pack_cfg(nl: int, taps: int, mlp_in: int, S: int, E: int, O: int) -> PackCfg:
    return PackCfg(n_layers=nl, taps=taps, mlp_in=mlp_in, S=S, E=E, O=O, Opad=CH * ((O + CH - 1) // CH))


def pack_offsets(nl: int, taps: int, mlp_in: int, S: int, E: int, O: int) -> List[int]:
    """Element offsets of the 8 packed segments (+ total) inside the buffer `pack_params` returns."""
    off = (C.c_longlong * 9)()
    cfg = _pack_cfg(nl, taps, mlp_in, S, E, O)
    if lib().gwn_pack_offsets(C.byref(cfg), off) < 0:
        check(-1, 'gwn_pack_offsets')
    return list(off)


@torch.library.custom_op('gwn::pack_params', mutates_args=())
def pack_params(params: List[Tensor], nl: int, taps: int, mlp_in: int, S: int, E: int, O: int) -> Tensor:
    """params = Wf[nl], bf[nl], Wg[nl], bg[nl], Wm[nl], Ws[nl], bs[nl], W1, W2, b2 (reference shapes, fp32) ->
    one flat fp32 buffer with the kernels' layouts (include/gwn.h: gwn_pack_params).  ONE launch."""
    for t in params:
        _req(t, torch.float32, 'parameter')
    cfg = _pack_cfg(nl, taps, mlp_in, S, E, O)
    ptrs = PackPtrs()
    for i, name in enumerate(('w_filter', 'b_filter', 'w_gate', 'b_gate', 'w_mlp', 'w_skip', 'b_skip')):
        arr = getattr(ptrs, name)
        for l in range(nl):
            arr[l] = params[i * nl + l].data_ptr()
    ptrs.w_end1, ptrs.w_end2, ptrs.b_end2 = (params[7 * nl + j].data_ptr() for j in range(3))
    out = torch.empty((pack_offsets(nl, taps, mlp_in, S, E, O)[8],), device=params[0].device, dtype=torch.float32)
    with torch.cuda.device(out.device):
        check(lib().gwn_pack_params(C.byref(cfg), C.byref(ptrs), _p(out), _stream()), 'gwn_pack_params')
    return out


@pack_params.register_fake
def _(params, nl, taps, mlp_in, S, E, O):
    Opad = CH * ((O + CH - 1) // CH)
    n = nl * taps * CH * 2 * CH + nl * 2 * CH + nl * mlp_in * CH + nl * CH * S + S + S * E + E * Opad + Opad
    return params[0].new_empty((n,))


def unpack_total(nl: int, taps: int, mlp_in: int, S: int, E: int, O: int) -> int:
    return nl * (2 * (CH * CH * taps + CH) + CH * mlp_in + CH * S + S) + E * S + O * E + O


@torch.library.custom_op('gwn::unpack_grads', mutates_args=())
def unpack_grads(g_wfg: List[Optional[Tensor]], g_bfg: List[Optional[Tensor]], g_wmlp: List[Optional[Tensor]],
                 g_wskip: Optional[Tensor], g_bskip: Optional[Tensor], g_wend1: Optional[Tensor],
                 g_wend2: Optional[Tensor], g_bend2: Optional[Tensor], taps: int, mlp_in: int, S: int, E: int,
                 O: int) -> Tensor:
    """Gradients of the packed tensors (None = absent) -> one flat fp32 buffer in parameter order (gwn_unpack_grads)."""
    nl = len(g_wfg)
    cfg = _pack_cfg(nl, taps, mlp_in, S, E, O)
    ptrs = UnpackPtrs()
    dev = None
    for name, lst in (('w_fg', g_wfg), ('b_fg', g_bfg), ('w_mlp', g_wmlp)):
        arr = getattr(ptrs, name)
        for l, t in enumerate(lst):
            if t is not None:
                _req(t, torch.float32, 'gradient'); dev = t.device
                arr[l] = t.data_ptr()
    for name, t in (('w_skip', g_wskip), ('b_skip', g_bskip), ('w_end1', g_wend1), ('w_end2', g_wend2), ('b_end2', g_bend2)):
        if t is not None:
            _req(t, torch.float32, 'gradient'); dev = t.device
            setattr(ptrs, name, t.data_ptr())
    if dev is None:
        raise ValueError('unpack_grads: no gradient given')
    out = torch.empty((unpack_total(nl, taps, mlp_in, S, E, O),), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(lib().gwn_unpack_grads(C.byref(cfg), C.byref(ptrs), _p(out), _stream()), 'gwn_unpack_grads')
    return out


@unpack_grads.register_fake
def _(g_wfg, g_bfg, g_wmlp, g_wskip, g_bskip, g_wend1, g_wend2, g_bend2, taps, mlp_in, S, E, O):
    ref = next(t for t in list(g_wfg) + list(g_bfg) + list(g_wmlp) + [g_wskip, g_bskip, g_wend1, g_wend2, g_bend2]
               if t is not None)
    return ref.new_empty((unpack_total(len(g_wfg), taps, mlp_in, S, E, O),), dtype=torch.float32)
