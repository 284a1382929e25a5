"""Drop-in `gwnet` nn.Module: same constructor, parameter/buffer names and forward contract as
``/root/reference/models/graph_wavenet.py:100-256``, with the whole block computed by the
hand-written sm_100a kernels behind ``torch.ops.gwn.*`` (see ops.py, include/gwn.h).

Two input conventions are accepted by ``forward``:
  * 3-D ``[num_nodes, horizon, in_dim]`` - the reference's literal call from ``Modified_UNET``
    (unet.py:224-226): the buffer is *reinterpreted* (not permuted) as ``[1, in_dim, V, horizon]``
    (graph_wavenet.py:189) and the result reinterpreted back to ``[V, horizon, out_dim]`` (:255);
  * 4-D ``[N, in_dim, V, T]`` - the general batched form used by every BASELINE config
    (the reference with its two hard-coded ``.view`` lines removed).

Activation precision: fp32 by default; bf16 storage (fp32 accumulate / statistics / gradients of
parameters) when ``compute_dtype = torch.bfloat16`` is set or the call runs under
``torch.autocast('cuda', dtype=torch.bfloat16)``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .ops import CH

_DEFAULT = object()


class nconv(nn.Module):
    """Node mixing ``einsum('ncvl,vw->ncwl')`` (graph_wavenet.py:60-66) on NCHW tensors, computed by
    the channels-last CUDA kernel (layout conversion on both sides; the fused block never uses this)."""

    def forward(self, x: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
        n, c, v, l = x.shape
        if c % CH != 0:
            raise ValueError(f'nconv kernel handles channel counts that are multiples of {CH}')
        xs = x.permute(0, 3, 2, 1).reshape(n * l, v, c // CH, CH).permute(2, 0, 1, 3).contiguous()
        ys = torch.stack([ops.node_mix(xs[i], A.contiguous().float(), False) for i in range(c // CH)])
        return ys.permute(1, 2, 0, 3).reshape(n, l, A.shape[1], c).permute(0, 3, 2, 1).contiguous()


class linear(nn.Module):
    """Parameter container for the gcn 1x1 'mlp' (graph_wavenet.py:68-74)."""

    def __init__(self, c_in: int, c_out: int):
        super().__init__()
        self.mlp = nn.Conv2d(c_in, c_out, kernel_size=(1, 1), padding=(0, 0), stride=(1, 1), bias=True)


class gcn(nn.Module):
    """Parameter container for the diffusion convolution (graph_wavenet.py:76-98)."""

    def __init__(self, c_in: int, c_out: int, dropout: float, support_len: int = 3, order: int = 2):
        super().__init__()
        self.nconv = nconv()
        self.mlp = linear((order * support_len + 1) * c_in, c_out)
        self.dropout = dropout
        self.order = order


def _pack_all(nl, p):
    """One gather launch: reference-shaped parameters -> the kernels' packed layouts (`gwn_pack_params`, csrc/pack.cu).
    p = Wf[nl], bf[nl], Wg[nl], bg[nl], Wm[nl], bm[nl], Ws[nl], bs[nl], W1, b1, W2, b2 (flat list).  Returns the views
    (w_fg [nl, k*32, 64], b_fg [nl, 64], w_mlp [nl, mlp_in, 32], w_skip [32*nl, S], b_skip [S], w_end1 [S, E],
    w_end2 [E, Opad], b_end2 [Opad]) of one flat buffer, and the dims."""
    Wf, bf, Wg, bg, Wm, bm, Ws, bs = (p[i * nl:(i + 1) * nl] for i in range(8))
    W1, b1, W2, b2 = p[8 * nl:]
    k, mlp_in, S = Wf[0].shape[3], Wm[0].shape[1], Ws[0].shape[0]
    O, E = W2.shape[0], W2.shape[1]
    Opad = CH * ((O + CH - 1) // CH)
    with torch.no_grad():
        srcs = [t.detach().contiguous() for t in (*Wf, *bf, *Wg, *bg, *Wm, *Ws, *bs, W1, W2, b2)]
        flat = ops.pack_params(srcs, nl, k, mlp_in, S, E, O)
    o = ops.pack_offsets(nl, k, mlp_in, S, E, O)
    views = (flat[o[0]:o[1]].view(nl, k * CH, 2 * CH), flat[o[1]:o[2]].view(nl, 2 * CH), flat[o[2]:o[3]].view(nl, mlp_in, CH),
             flat[o[3]:o[4]].view(nl * CH, S), flat[o[4]:o[5]], flat[o[5]:o[6]].view(S, E), flat[o[6]:o[7]].view(E, Opad),
             flat[o[7]:o[8]])
    return views, (nl, k, S, E, O, mlp_in)


def _unpacked_views(flat, dims):
    """Parameter-shaped views of the flat buffer `gwn_unpack_grads` writes (parameter order, csrc/pack.cu)."""
    nl, k, S, E, O, mlp_in = dims
    nw = CH * CH * k
    per = 2 * (nw + CH) + CH * mlp_in + CH * S + S
    dWf, dbf, dWg, dbg, dWm, dWs, dbs = ([None] * nl for _ in range(7))
    for l in range(nl):
        b0 = l * per
        dWf[l] = flat[b0:b0 + nw].view(CH, CH, 1, k)
        dbf[l] = flat[b0 + nw:b0 + nw + CH]
        dWg[l] = flat[b0 + nw + CH:b0 + 2 * nw + CH].view(CH, CH, 1, k)
        dbg[l] = flat[b0 + 2 * nw + CH:b0 + 2 * (nw + CH)]
        b1 = b0 + 2 * (nw + CH)
        dWm[l] = flat[b1:b1 + CH * mlp_in].view(CH, mlp_in, 1, 1)
        b2 = b1 + CH * mlp_in
        dWs[l] = flat[b2:b2 + CH * S].view(S, CH, 1, 1)
        dbs[l] = flat[b2 + CH * S:b2 + CH * S + S]
    t0 = nl * per
    dW1 = flat[t0:t0 + E * S].view(E, S, 1, 1)
    dW2 = flat[t0 + E * S:t0 + E * S + O * E].view(O, E, 1, 1)
    db2 = flat[t0 + E * S + O * E:t0 + E * S + O * E + O]
    return dWf, dbf, dWg, dbg, dWm, dWs, dbs, dW1, dW2, db2


class _PackLayers(torch.autograd.Function):
    """Autograd edge from the per-layer parameters (filter / gate convs, gcn mlp weight) to their packed views; backward
    scatters the packed gradients back with ONE launch (`gwn_unpack_grads`).  The packing itself happened in `_pack_all`.

    inputs : dims, w_fg [nl,..], b_fg [nl,..], w_mlp [nl,..] (packed views), Wf[nl], bf[nl], Wg[nl], bg[nl], Wm[nl]
    outputs: w_fg[nl], b_fg[nl], w_mlp[nl]"""

    @staticmethod
    def forward(ctx, dims, w_fg, b_fg, w_mlp, *p):
        ctx.dims = dims
        ctx.set_materialize_grads(False)
        return (*w_fg.unbind(0), *b_fg.unbind(0), *w_mlp.unbind(0))

    @staticmethod
    def backward(ctx, *g):
        nl, k, S, E, O, mlp_in = ctx.dims
        cont = lambda t: None if t is None else t.contiguous()  # noqa: E731
        g_wfg, g_bfg, g_wmlp = [cont(t) for t in g[:nl]], [cont(t) for t in g[nl:2 * nl]], [cont(t) for t in g[2 * nl:3 * nl]]
        if all(t is None for t in (*g_wfg, *g_bfg, *g_wmlp)):
            return (None,) * (4 + 5 * nl)
        flat = ops.unpack_grads(g_wfg, g_bfg, g_wmlp, None, None, None, None, None, k, mlp_in, S, E, O)
        dWf, dbf, dWg, dbg, dWm, _dWs, _dbs, _dW1, _dW2, _db2 = _unpacked_views(flat, ctx.dims)
        for l in range(nl):                                    # parameters the block never used keep grad None
            if g_wfg[l] is None:
                dWf[l] = dWg[l] = None
            if g_bfg[l] is None:
                dbf[l] = dbg[l] = None
            if g_wmlp[l] is None:
                dWm[l] = None
        return (None, None, None, None, *dWf, *dbf, *dWg, *dbg, *dWm)


class _PackHead(torch.autograd.Function):
    """Same for the head's parameters (skip convs, end_conv_1 weight, end_conv_2).  Applied right before the head in
    forward, so that in backward it runs right AFTER the head (autograd runs the youngest ready node first): the head's
    parameter gradients - two thirds of all gradient bytes - are final at the very start of backward and their
    data-parallel exchange (ddp.BucketedGradAllReduce, bucket 0) overlaps the backward of every layer.

    inputs : dims, w_skip, b_skip, w_end1, w_end2, b_end2 (packed views), Ws[nl], bs[nl], W1, W2, b2
    outputs: w_skip, b_skip, w_end1, w_end2, b_end2"""

    @staticmethod
    def forward(ctx, dims, w_skip, b_skip, w_end1, w_end2, b_end2, *p):
        ctx.dims = dims
        ctx.set_materialize_grads(False)
        return tuple(t.view_as(t) for t in (w_skip, b_skip, w_end1, w_end2, b_end2))

    @staticmethod
    def backward(ctx, *g):
        nl, k, S, E, O, mlp_in = ctx.dims
        g_wskip, g_bskip, g_wend1, g_wend2, g_bend2 = (None if t is None else t.contiguous() for t in g)
        if all(t is None for t in (g_wskip, g_bskip, g_wend1, g_wend2, g_bend2)):
            return (None,) * (6 + 2 * nl + 3)
        none = [None] * nl
        flat = ops.unpack_grads(none, none, none, g_wskip, g_bskip, g_wend1, g_wend2, g_bend2, k, mlp_in, S, E, O)
        _a, _b, _c, _d, _e, dWs, dbs, dW1, dW2, db2 = _unpacked_views(flat, ctx.dims)
        if g_wskip is None:
            dWs = none
        if g_bskip is None:
            dbs = none
        return (None, None, None, None, None, None, *dWs, *dbs, dW1 if g_wend1 is not None else None,
                dW2 if g_wend2 is not None else None, db2 if g_bend2 is not None else None)


class gwnet(nn.Module):
    _instances = 0          # numbers the modules of a process (fused dropout stream key)

    def __init__(self, device, num_nodes=67, dropout=0.3, supports=_DEFAULT, gcn_bool=True, addaptadj=True,
                 aptinit=None, in_dim=256, out_dim=255, horizon=1, residual_channels=32, dilation_channels=32,
                 skip_channels=256, end_channels=512, kernel_size=1, blocks=4, layers=2):
        super().__init__()
        if residual_channels != CH or dilation_channels != CH:
            raise NotImplementedError(
                f'the sm_100a kernels are specialised for residual_channels == dilation_channels == {CH} '
                '(the value every reference configuration uses)')
        if skip_channels % CH or end_channels % CH:
            raise NotImplementedError('skip_channels and end_channels must be multiples of 32')
        self.dropout = dropout
        self.blocks = blocks
        self.layers = layers
        self.gcn_bool = gcn_bool
        self.addaptadj = addaptadj
        self.horizon = horizon
        self.num_nodes = num_nodes
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.kernel_size = kernel_size
        self.compute_dtype: Optional[torch.dtype] = None    # None: follow autocast, else fp32
        self.use_tensor_cores = True                        # bf16 hops on tcgen05 when the supports fit on chip
        self.dropout_mode = 'fused'                         # 'fused' (in-kernel Philox) | 'torch' (F.dropout mask)
        self.sparse_supports = True                         # V > 80: fixed supports with few non-zeros per row run as sparse gathers

        # registration order below mirrors graph_wavenet.py:110-183 so state_dict keys (and a seeded
        # default init) line up with the reference
        self.filter_convs = nn.ModuleList()
        self.gate_convs = nn.ModuleList()
        self.residual_convs = nn.ModuleList()
        self.skip_convs = nn.ModuleList()
        self.bn = nn.ModuleList()
        self.gconv = nn.ModuleList()
        self.start_conv = nn.Conv2d(in_dim, residual_channels, kernel_size=(1, 1))

        if supports is _DEFAULT:
            # what the reference's own load_adj('doubletransition') yields: one identity (graph_wavenet.py:24,51)
            supports = [torch.eye(num_nodes, dtype=torch.float32)]
        self.supports_len = 0 if supports is None else len(supports)
        self.supports: List[torch.Tensor] = [] if supports is None else \
            [torch.as_tensor(s, dtype=torch.float32).to(device).contiguous() for s in supports]

        if gcn_bool and addaptadj:
            if aptinit is None:
                self.nodevec1 = nn.Parameter(torch.randn(num_nodes, 10).to(device), requires_grad=True)
                self.nodevec2 = nn.Parameter(torch.randn(10, num_nodes).to(device), requires_grad=True)
            else:
                m, p, n = torch.svd(torch.as_tensor(aptinit, dtype=torch.float32))
                self.nodevec1 = nn.Parameter(torch.mm(m[:, :10], torch.diag(p[:10] ** 0.5)).to(device))
                self.nodevec2 = nn.Parameter(torch.mm(torch.diag(p[:10] ** 0.5), n[:, :10].t()).to(device))
            self.supports_len += 1

        receptive_field = 1
        self.dilations: List[int] = []
        for _b in range(blocks):
            additional_scope = kernel_size - 1
            new_dilation = 1
            for _i in range(layers):
                self.filter_convs.append(nn.Conv2d(residual_channels, dilation_channels, (1, kernel_size),
                                                   dilation=new_dilation))
                self.gate_convs.append(nn.Conv2d(residual_channels, dilation_channels, (1, kernel_size),
                                                 dilation=new_dilation))
                self.residual_convs.append(nn.Conv2d(dilation_channels, residual_channels, (1, 1)))
                self.skip_convs.append(nn.Conv2d(dilation_channels, skip_channels, (1, 1)))
                self.bn.append(nn.BatchNorm2d(residual_channels))
                self.dilations.append(new_dilation)
                new_dilation *= 2
                receptive_field += additional_scope
                additional_scope *= 2
                if gcn_bool:
                    self.gconv.append(gcn(dilation_channels, residual_channels, dropout,
                                          support_len=self.supports_len))
        self.end_conv_1 = nn.Conv2d(skip_channels, end_channels, (1, 1), bias=True)
        self.end_conv_2 = nn.Conv2d(end_channels, out_dim, (1, 1), bias=True)
        self.receptive_field = receptive_field
        self._sparse_checked = False
        self._rng_state: Optional[torch.Tensor] = None
        self._instance = gwnet._instances
        gwnet._instances += 1
        self._calls = 0
        self.to(device)

    # supports are plain attributes in the reference (not buffers, not in the state_dict); keep that, but
    # let .to()/.cuda() move them with the module
    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        self.supports = [fn(s).contiguous() for s in self.supports]
        self._sparse_checked = False
        if self._rng_state is not None:
            self._rng_state = fn(self._rng_state)
        return self

    # ------------------------------------------------------------------ helpers
    def _act_dtype(self) -> torch.dtype:
        if self.compute_dtype is not None:
            return self.compute_dtype
        if torch.is_autocast_enabled('cuda') and torch.get_autocast_dtype('cuda') == torch.bfloat16:
            return torch.bfloat16
        return torch.float32

    def layer_lengths(self, t_in: int) -> List[int]:
        L = [max(t_in, self.receptive_field)]
        for d in self.dilations:
            L.append(L[-1] - d * (self.kernel_size - 1))
        return L

    def _packed(self):
        """Packs every parameter with one launch; returns the layer views (autograd-connected) and a closure that
        connects the head's views - to be called right before the head (see _PackHead)."""
        nl = self.blocks * self.layers
        mlps = [g.mlp.mlp for g in self.gconv] if self.gcn_bool else list(self.residual_convs)
        Wf, bf = [c.weight for c in self.filter_convs], [c.bias for c in self.filter_convs]
        Wg, bg = [c.weight for c in self.gate_convs], [c.bias for c in self.gate_convs]
        Wm, bm = [m.weight for m in mlps], [m.bias for m in mlps]
        Ws, bs = [c.weight for c in self.skip_convs], [c.bias for c in self.skip_convs]
        head = [self.end_conv_1.weight, self.end_conv_1.bias, self.end_conv_2.weight, self.end_conv_2.bias]
        (w_fg, b_fg, w_mlp, w_skip, b_skip, w_end1, w_end2, b_end2), dims = _pack_all(nl, Wf + bf + Wg + bg + Wm + bm + Ws + bs + head)
        out = _PackLayers.apply(dims, w_fg, b_fg, w_mlp, *Wf, *bf, *Wg, *bg, *Wm)

        def head_views():
            h = _PackHead.apply(dims, w_skip, b_skip, w_end1, w_end2, b_end2, *Ws, *bs, head[0], head[2], head[3])
            return dict(w_skip=h[0], b_skip=h[1], w_end1=h[2], w_end2=h[3], b_end2=h[4])
        return dict(w_fg=out[:nl], b_fg=out[nl:2 * nl], w_mlp=out[2 * nl:3 * nl], b_mlp=bm, head=head_views)

    # ------------------------------------------------------------------ forward
    def forward(self, input: torch.Tensor, dropout_masks: Optional[Sequence[Optional[torch.Tensor]]] = None):
        literal = input.dim() == 3
        if literal:                                          # graph_wavenet.py:189
            x = input.reshape(1, self.in_dim, self.num_nodes, self.horizon)
        else:
            x = input
        y = self._forward_nchw(x, dropout_masks)
        if literal:                                          # graph_wavenet.py:255
            return y.reshape(self.num_nodes, self.horizon, -1)
        return y

    def _forward_nchw(self, x, dropout_masks):
        if not x.is_cuda:
            raise ops._lib.GwnError('gwnet runs on a CUDA (B200) device only - there is no CPU fallback')
        N, Cin, V, T = x.shape
        if Cin != self.in_dim or V != self.num_nodes:
            raise ValueError(f'expected input [N,{self.in_dim},{self.num_nodes},T], got {tuple(x.shape)}')
        dt = self._act_dtype()
        nl = self.blocks * self.layers
        L = self.layer_lengths(T)
        Lf = L[-1]
        if Lf < 1:
            raise ValueError(f'input length {T} too short for receptive field {self.receptive_field}')
        training = self.training
        pk = self._packed()

        supports: List[torch.Tensor] = []
        if self.gcn_bool:
            supports = list(self.supports)
            if not self._sparse_checked and V > 80 and self.sparse_supports:
                # big graphs: fixed supports with few neighbours per node are applied as sparse gathers (ops.register_sparse_support)
                for s_ in supports:
                    if s_.is_cuda:
                        ops.register_sparse_support(s_)
                self._sparse_checked = True
            if self.addaptadj:                               # graph_wavenet.py:201-203
                # bf16 path with the supports resident on chip: the adjacency travels as a pair [2,V,V] so the fused
                # backward can return its gradient in factored form (ops.AdaptiveAdjacency)
                pair = (dt == torch.bfloat16 and self.use_tensor_cores and training and
                        ops.hop_mode(V, len(supports) + 1) == 1)
                supports = supports + [ops.AdaptiveAdjacency.apply(self.nodevec1, self.nodevec2, pair)]

        p_drop = float(self.dropout) if (training and self.gcn_bool) else 0.0
        rng = None
        masks: List[Optional[torch.Tensor]] = [None] * nl
        if p_drop > 0.0:
            if dropout_masks is not None:                    # explicit NCHW masks (parity tests)
                masks = [None if m is None else m.permute(0, 3, 2, 1).contiguous().to(dt) for m in dropout_masks]
            elif self.dropout_mode == 'torch':               # the reference's own Philox draws (graph_wavenet.py:97)
                masks = [torch.nn.functional.dropout(x.new_ones((N, CH, V, L[i + 1])), p_drop, True)
                         .permute(0, 3, 2, 1).contiguous().to(dt) for i in range(nl)]
            else:
                if self._rng_state is None or self._rng_state.device != x.device:
                    # key of the fused dropout stream: torch's seed mixed with the data-parallel rank (every rank draws
                    # its own masks for its own shard) and a per-instance number (two gwnet modules in one process do not
                    # share a stream).  Not a registered buffer: the reference's state_dict has no such entry.
                    rank = torch.distributed.get_rank() if (torch.distributed.is_available() and
                                                            torch.distributed.is_initialized()) else 0
                    seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * (rank + 1) + 0xD1B54A32D192ED03 * self._instance)
                    self._rng_state = torch.tensor([seed & (2 ** 62 - 1), 0], dtype=torch.int64, device=x.device)
                rng = self._rng_state.clone()                # this step's {seed, offset}
                self._rng_state[1] += nl                     # graph-safe: advances on every replay

        # bf16 + supports that fit on chip: diffusion hops run on the tcgen05 tensor cores; the UMMA operand
        # images of (A, A^2, A^T, (A^2)^T) are built once per forward and shared by all layers
        hop_mats = None
        if dt == torch.bfloat16 and supports and self.use_tensor_cores:
            sup_c = [(s[0] if s.dim() == 3 else s).detach().contiguous() for s in supports]
            # V <= 80: every support image stays resident in shared memory; larger graphs (the 3,100-node
            # configurations): one TMA-tiled tensor-core GEMM per hop
            # (ops.hop_mode = the C library's own rule: both sides agree on which images the buffer holds)
            hop_mats = ops.hop_mats(sup_c) if ops.hop_mode(V, len(sup_c)) == 1 else ops.support_images(sup_c)

        u = ops.StartConv.apply(x, self.start_conv.weight, self.start_conv.bias, L[0], dt == torch.bfloat16)
        # per-step gradient workspace (bf16 tensor-core path): ONE zero-filled buffer whose slices receive every layer's
        # small accumulated gradients (BN-backward statistics, dW, db) and ONE accumulator for the adaptive support's
        # gradient that all layers add into - instead of a fill per layer and an autograd add per layer
        grad_ws = None
        if (training and torch.is_grad_enabled() and hop_mats is not None and dt == torch.bfloat16 and self.gcn_bool and
                self.kernel_size <= 4 and sum(bool(s.requires_grad) for s in supports) <= 1):
            mlp_in = CH * (1 + 2 * len(supports))
            per = sum(ops._layer_bwd_sizes(self.kernel_size, mlp_in, V, [], True))
            per = (per + 3) // 4 * 4
            acc_n = sum(s.numel() for s in supports if s.requires_grad)
            ws = torch.zeros(nl * per + acc_n, device=x.device, dtype=torch.float32)
            d_acc = None
            for s in supports:
                if s.requires_grad:
                    d_acc = ws[nl * per:].view(s.shape)
            grad_ws = [dict(flat=ws[i * per:(i + 1) * per], d_acc=d_acc, first=(i == 0)) for i in range(nl)]
        stats = None
        z_last = []
        for i in range(nl):
            last = i == nl - 1
            bn_prev = self.bn[i - 1] if i > 0 else None
            has_gconv = (not last) or training               # last layer's gconv/bn only feed running stats
            meta = dict(training=training, momentum=0.1 if bn_prev is None or bn_prev.momentum is None
                        else bn_prev.momentum, eps=1e-5 if bn_prev is None else bn_prev.eps, Lf=Lf,
                        taps=self.kernel_size, dilation=self.dilations[i], order=2, has_gconv=has_gconv,
                        dropout_p=p_drop if masks[i] is None else 1.0, seed=0, offset=i)
            if masks[i] is not None:
                meta['dropout_p'] = p_drop
            if grad_ws is not None:
                meta['grad_ws'] = grad_ws[i]
            u, stats, zl = ops.WaveNetLayer.apply(
                u, stats,
                None if bn_prev is None else bn_prev.weight, None if bn_prev is None else bn_prev.bias,
                None if bn_prev is None else bn_prev.running_mean, None if bn_prev is None else bn_prev.running_var,
                pk['w_fg'][i], pk['b_fg'][i], pk['w_mlp'][i] if has_gconv else None,
                pk['b_mlp'][i] if has_gconv else None, masks[i], rng, hop_mats, meta, *supports)
            z_last.append(zl)
        if training:
            # the reference still runs bn[last] (its output is dead, :250-252) - keep its running stats in step
            with torch.no_grad():
                bl = self.bn[nl - 1]
                ops.bn_fold(stats, float(N * L[nl] * V), bl.weight, bl.bias, bl.running_mean, bl.running_var,
                            0.1 if bl.momentum is None else bl.momentum, bl.eps, True)
                torch._foreach_add_([b.num_batches_tracked for b in self.bn], 1)
        hk = pk['head']()
        return ops.SkipHead.apply(hk['w_skip'], hk['b_skip'], hk['w_end1'], self.end_conv_1.bias, hk['w_end2'],
                                  hk['b_end2'], self.out_dim, *z_last)
