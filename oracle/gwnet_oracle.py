"""CPU restatement of the reference Graph WaveNet block.  TEST INFRASTRUCTURE ONLY.

A from-scratch, purely functional restatement (explicit tap sums, explicit
batch-norm arithmetic, explicit hop recursion) of what
``/root/reference/models/graph_wavenet.py:60-256`` computes.  It is deliberately
written *differently* from the reference (no ``nn.Module``, no ``conv2d``) so that
agreeing with the reference-generated golden vectors (``tests/golden``) is a real
check and not a tautology.  Works in fp32 or fp64 (dtype follows the inputs), on
CPU (tests also run it on ``cuda`` as a convenience; never on the product path).

Layout here is the reference's: activations are ``[N, C, V, L]`` (batch, channel,
node, time), time innermost.

Parity: pinned against reference-executed golden vectors, see ``oracle/__init__``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

# When True the 1x1 / dilated convolutions and BatchNorm go through the same ATen library calls the
# reference's nn.Conv2d / nn.BatchNorm2d make (oneDNN on CPU).  Same arithmetic, and the honest form
# to TIME as the reference's CPU path (bench.py cpu_baseline / --impl reference); the explicit
# tap-sum form below stays the default for parity checks.  Both are pinned by the golden tests.
ATEN_PATH = False


# --------------------------------------------------------------------------- storage-rounding emulation
class _StoreRound(torch.autograd.Function):
    """Emulates a tensor that the 16-bit data path STORES in `dtype`: the forward value is rounded to
    `dtype` (then carried on in the oracle's working precision) and so is the gradient flowing back
    through it (gradient activations are stored in the same 16-bit type)."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.dtype = dtype
        return x.to(dtype).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype).to(g.dtype), None


class _GateSavedRounded(torch.autograd.Function):
    """z = tanh(f)*sigmoid(g) where the backward uses the 16-bit-rounded saved tanh/sigmoid values."""

    @staticmethod
    def forward(ctx, f, g, dtype):
        a, b = torch.tanh(f), torch.sigmoid(g)
        ctx.save_for_backward(a.to(dtype).to(a.dtype), b.to(dtype).to(b.dtype))
        return a * b

    @staticmethod
    def backward(ctx, dz):
        a, b = ctx.saved_tensors
        return dz * b * (1 - a * a), dz * a * b * (1 - b), None


# tensors the storage emulation leaves unrounded (names: 'u' = residual stream and its gradient, 'z', 'hops',
# 'gate_saved'); set by tests / scripts that ask what keeping one of them in fp32 would buy
STORAGE_EXCLUDE: set = set()


def _sr(x, dtype, what=None):
    if dtype is None or (what is not None and what in STORAGE_EXCLUDE):
        return x
    return _StoreRound.apply(x, dtype)


# --------------------------------------------------------------------------- config
@dataclass
class GWNetConfig:
    """Constructor arguments of the reference ``gwnet`` (graph_wavenet.py:101)."""
    num_nodes: int = 67
    in_dim: int = 2
    out_dim: int = 12
    residual_channels: int = 32
    dilation_channels: int = 32
    skip_channels: int = 256
    end_channels: int = 512
    kernel_size: int = 2
    blocks: int = 4
    layers: int = 2
    n_fixed_supports: int = 2       # len(supports) handed to the ctor
    adaptive: bool = True           # addaptadj (only meaningful when gcn_bool)
    gcn_bool: bool = True           # False -> residual_convs path (graph_wavenet.py:245)
    order: int = 2                  # gcn default (graph_wavenet.py:77)
    dropout: float = 0.3
    bn_eps: float = 1e-5            # nn.BatchNorm2d defaults (graph_wavenet.py:167)
    bn_momentum: float = 0.1

    @property
    def n_layers(self) -> int:
        return self.blocks * self.layers

    @property
    def n_supports(self) -> int:
        return self.n_fixed_supports + (1 if self.adaptive else 0)


def dilation_schedule(cfg: GWNetConfig) -> List[int]:
    """Dilation of layer i: doubles inside a block, resets per block
    (graph_wavenet.py:145-148,168)."""
    out = []
    for _ in range(cfg.blocks):
        d = 1
        for _ in range(cfg.layers):
            out.append(d)
            d *= 2
    return out


def receptive_field(cfg: GWNetConfig) -> int:
    """graph_wavenet.py:122,146,169-170,185."""
    rf = 1
    for _ in range(cfg.blocks):
        scope = cfg.kernel_size - 1
        for _ in range(cfg.layers):
            rf += scope
            scope *= 2
    return rf


def layer_lengths(cfg: GWNetConfig, t_in: int) -> List[int]:
    """[L0, L1, ...]: time length entering layer 0 and leaving each layer
    (pad to rf at graph_wavenet.py:191-195; conv shrink d*(k-1) per layer)."""
    L = [max(t_in, receptive_field(cfg))]
    for d in dilation_schedule(cfg):
        L.append(L[-1] - d * (cfg.kernel_size - 1))
    return L


# --------------------------------------------------------------------------- pieces
def adaptive_adjacency(e1: torch.Tensor, e2: torch.Tensor) -> torch.Tensor:
    """softmax(relu(E1 @ E2), dim=1)  (graph_wavenet.py:202)."""
    m = e1 @ e2
    m = torch.where(m > 0, m, torch.zeros_like(m))
    m = m - m.max(dim=1, keepdim=True).values
    e = torch.exp(m)
    return e / e.sum(dim=1, keepdim=True)


def node_mix(x: torch.Tensor, a: torch.Tensor) -> torch.Tensor:
    """y[n,c,w,l] = sum_v x[n,c,v,l] * A[v,w]   (nconv, graph_wavenet.py:65)."""
    if ATEN_PATH:
        return torch.einsum('ncvl,vw->ncwl', x, a).contiguous()
    n, c, v, l = x.shape
    xt = x.permute(0, 1, 3, 2).reshape(-1, v)          # rows (n,c,l), cols v
    return (xt @ a).reshape(n, c, l, a.shape[1]).permute(0, 1, 3, 2)


def pointwise(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """1x1 Conv2d: w is [O, C, 1, 1]  (graph_wavenet.py:71,117,164,174,179)."""
    if ATEN_PATH:
        return F.conv2d(x, w, b)
    y = torch.einsum('oc,ncvl->novl', w[:, :, 0, 0], x)
    if b is not None:
        y = y + b.view(1, -1, 1, 1)
    return y


def dilated_conv(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, d: int) -> torch.Tensor:
    """Conv2d kernel (1,k), dilation d, no padding (graph_wavenet.py:150-156):
    y[n,o,v,t] = b[o] + sum_{c,j} w[o,c,0,j] * x[n,c,v,t+j*d]."""
    if ATEN_PATH:
        return F.conv2d(x, w, b, dilation=d)
    k = w.shape[3]
    lout = x.shape[3] - d * (k - 1)
    y = None
    for j in range(k):
        term = torch.einsum('oc,ncvl->novl', w[:, :, 0, j], x[:, :, :, j * d:j * d + lout])
        y = term if y is None else y + term
    return y + b.view(1, -1, 1, 1)


def diffusion_conv(z: torch.Tensor, supports: Sequence[torch.Tensor], w: torch.Tensor,
                   b: torch.Tensor, order: int, storage=None) -> torch.Tensor:
    """gcn.forward before dropout (graph_wavenet.py:85-96): concat order is
    [z, zA0, zA0^2, zA1, zA1^2, ...]; powers are sequential re-applications."""
    pieces = [z]
    for a in supports:
        y = z
        for _ in range(order):
            y = _sr(node_mix(y, a), storage, 'hops')
            pieces.append(y)
    return pointwise(torch.cat(pieces, dim=1), w, b)


def batch_norm_train(u: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float):
    """Training-mode BatchNorm2d (graph_wavenet.py:167,250): biased variance
    normalises; returns (y, batch_mean, biased_var)."""
    if ATEN_PATH:
        with torch.no_grad():
            mean = u.mean(dim=(0, 2, 3))
            var = u.var(dim=(0, 2, 3), unbiased=False)
        return F.batch_norm(u, None, None, gamma, beta, True, 0.0, eps), mean, var
    mean = u.mean(dim=(0, 2, 3))
    var = ((u - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
    y = (u - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + eps)
    return y * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1), mean, var


def batch_norm_eval(u, gamma, beta, rmean, rvar, eps):
    y = (u - rmean.view(1, -1, 1, 1)) / torch.sqrt(rvar.view(1, -1, 1, 1) + eps)
    return y * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


# --------------------------------------------------------------------------- whole block
@dataclass
class ForwardTrace:
    """Intermediates kept for per-kernel parity tests."""
    adp: Optional[torch.Tensor] = None
    x_in: List[torch.Tensor] = field(default_factory=list)    # layer inputs (post-BN)
    z: List[torch.Tensor] = field(default_factory=list)       # gate outputs
    h: List[torch.Tensor] = field(default_factory=list)       # gcn outputs (after dropout mask)
    u: List[torch.Tensor] = field(default_factory=list)       # pre-BN (h + residual)
    skip: Optional[torch.Tensor] = None
    new_running: Dict[str, torch.Tensor] = field(default_factory=dict)


def gwnet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor,
                  fixed_supports: Sequence[torch.Tensor], cfg: GWNetConfig, *,
                  training: bool = True,
                  dropout_masks: Optional[Sequence[Optional[torch.Tensor]]] = None,
                  trace: Optional[ForwardTrace] = None, storage=None, head_masks=None) -> torch.Tensor:
    """General-mode ``gwnet.forward`` (graph_wavenet.py:188-256 without the two
    literal ``.view`` statements at :189 and :255).

    ``sd`` maps the reference's state_dict names to tensors (leaf tensors with
    ``requires_grad`` give gradients by autograd).  ``x`` is ``[N, in_dim, V, T]``.
    ``dropout_masks[i]`` (optional, training only) is the multiplicative mask
    (0 or 1/(1-p)) applied to layer i's gcn output, ``[N, C, V, L_i]``; ``None``
    means dropout is inactive (p = 0 or eval).  Running statistics are not
    mutated; their would-be new values are returned through ``trace``.

    ``storage`` (None or torch.bfloat16): emulate the 16-bit data path by rounding every tensor that
    path STORES in 16 bits (the pre-BN stream u, the gate output z, every diffusion hop, the saved
    tanh/sigmoid, and the gradients flowing through them), at the points the CUDA kernels store them;
    all arithmetic between those points stays in the oracle's working precision (the kernels
    accumulate in fp32).  Used for the bf16 parity tests: ReLU/relu-mask decisions then see the same
    rounded activations as the kernels do ("bf16 parity").

    ``head_masks`` (None or (m1 [N,S,V,Lf], m2 [N,E,V,Lf]) of 0/1): replace the two ReLUs of the head by a
    multiplication with the given masks (the ReLU DECISIONS of another evaluation).  A ReLU network is piecewise
    linear in its activations; with the decisions pinned, output and gradients are smooth in the remaining
    rounding noise, which is what a 2e-2 gradient bar can be asked of (tests/test_gpu_parity.py).
    """
    rf = receptive_field(cfg)
    t_in = x.shape[3]
    if t_in < rf:                                            # :191-195
        x = torch.cat([x.new_zeros(x.shape[0], x.shape[1], x.shape[2], rf - t_in), x], dim=3)
    h = _sr(pointwise(x, sd['start_conv.weight'], sd['start_conv.bias']), storage, 'u')      # :196

    supports = list(fixed_supports)
    if cfg.gcn_bool and cfg.adaptive:                        # :201-203
        adp = adaptive_adjacency(sd['nodevec1'], sd['nodevec2'])
        supports = supports + [adp]
        if trace is not None:
            trace.adp = adp

    dil = dilation_schedule(cfg)
    skip = None
    for i in range(cfg.n_layers):                            # :206
        res = h
        if trace is not None:
            trace.x_in.append(res)
        f = dilated_conv(res, sd[f'filter_convs.{i}.weight'], sd[f'filter_convs.{i}.bias'], dil[i])
        g = dilated_conv(res, sd[f'gate_convs.{i}.weight'], sd[f'gate_convs.{i}.bias'], dil[i])
        if storage is None:
            z = torch.tanh(f) * torch.sigmoid(g)             # :222-226
        else:
            z = _sr(torch.tanh(f) * torch.sigmoid(g) if 'gate_saved' in STORAGE_EXCLUDE
                    else _GateSavedRounded.apply(f, g, storage), storage, 'z')
        s = pointwise(z, sd[f'skip_convs.{i}.weight'], sd[f'skip_convs.{i}.bias'])   # :231
        skip = s if skip is None else s + skip[:, :, :, -s.shape[3]:]                # :232-236
        if cfg.gcn_bool:
            hh = diffusion_conv(z, supports, sd[f'gconv.{i}.mlp.mlp.weight'],
                                sd[f'gconv.{i}.mlp.mlp.bias'], cfg.order, storage)   # :241
            if training and dropout_masks is not None and dropout_masks[i] is not None:
                hh = hh * dropout_masks[i]                   # :97
        else:
            hh = pointwise(z, sd[f'residual_convs.{i}.weight'], sd[f'residual_convs.{i}.bias'])  # :245
        u = hh + res[:, :, :, -hh.shape[3]:]                 # :247
        if training:                                         # :250
            if storage is None:
                h, mean, var = batch_norm_train(u, sd[f'bn.{i}.weight'], sd[f'bn.{i}.bias'], cfg.bn_eps)
            else:   # statistics from the fp32 epilogue values, normalisation applied to the stored (rounded) u
                mean = u.mean(dim=(0, 2, 3))
                var = ((u - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
                us = _sr(u, storage, 'u')
                h = (us - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + cfg.bn_eps)
                h = h * sd[f'bn.{i}.weight'].view(1, -1, 1, 1) + sd[f'bn.{i}.bias'].view(1, -1, 1, 1)
            if trace is not None:
                cnt = u.numel() // u.shape[1]
                unbiased = var * (cnt / max(cnt - 1, 1))
                m = cfg.bn_momentum
                trace.new_running[f'bn.{i}.running_mean'] = \
                    (1 - m) * sd[f'bn.{i}.running_mean'] + m * mean.detach()
                trace.new_running[f'bn.{i}.running_var'] = \
                    (1 - m) * sd[f'bn.{i}.running_var'] + m * unbiased.detach()
        else:
            h = batch_norm_eval(_sr(u, storage, 'u'), sd[f'bn.{i}.weight'], sd[f'bn.{i}.bias'],
                                sd[f'bn.{i}.running_mean'], sd[f'bn.{i}.running_var'], cfg.bn_eps)
        if trace is not None:
            trace.z.append(z)
            trace.h.append(hh)
            trace.u.append(u)
    if trace is not None:
        trace.skip = skip
    if head_masks is not None:
        y = skip * head_masks[0].to(skip.dtype)
        y = pointwise(y, sd['end_conv_1.weight'], sd['end_conv_1.bias']) * head_masks[1].to(skip.dtype)
        return pointwise(y, sd['end_conv_2.weight'], sd['end_conv_2.bias'])
    y = torch.relu(skip)                                     # :252
    y = torch.relu(pointwise(y, sd['end_conv_1.weight'], sd['end_conv_1.bias']))   # :253
    return pointwise(y, sd['end_conv_2.weight'], sd['end_conv_2.bias'])            # :254


def gwnet_forward_literal(sd, inp, fixed_supports, cfg: GWNetConfig, horizon: int, **kw):
    """Literal-mode forward: the two reinterpreting ``.view``s of the reference
    (graph_wavenet.py:189, :255): ``[67,h,F] -> view(1,F,67,h)`` and back."""
    f = cfg.in_dim
    x = inp.reshape(1, f, cfg.num_nodes, horizon)            # memory reinterpretation, not a permute
    y = gwnet_forward(sd, x, fixed_supports, cfg, **kw)
    return y.reshape(cfg.num_nodes, horizon, cfg.out_dim)


# --------------------------------------------------------------------------- closed-form backward pieces
def adaptive_adjacency_backward(e1, e2, grad_p):
    """dR = P*(dP - rowsum(dP*P)); dM = dR*[M>0]; dE1 = dM E2^T; dE2 = E1^T dM."""
    m = e1 @ e2
    p = adaptive_adjacency(e1, e2)
    dr = p * (grad_p - (grad_p * p).sum(dim=1, keepdim=True))
    dm = dr * (m > 0).to(dr.dtype)
    return dm @ e2.t(), e1.t() @ dm


def gate_backward(f, g, dz):
    """df = dz*sig(g)*(1-tanh(f)^2); dg = dz*tanh(f)*sig(g)*(1-sig(g))."""
    a, b = torch.tanh(f), torch.sigmoid(g)
    return dz * b * (1 - a * a), dz * a * b * (1 - b)


def node_mix_backward(x, a, gy):
    """nconv backward: dx = einsum('ncwl,vw->ncvl'), dA = einsum('ncvl,ncwl->vw')."""
    return torch.einsum('ncwl,vw->ncvl', gy, a), torch.einsum('ncvl,ncwl->vw', x, gy)


# --------------------------------------------------------------------------- deterministic synthetic parameters
def state_dict_shapes(cfg: GWNetConfig) -> Dict[str, tuple]:
    """Names/shapes of the reference state_dict, in the
    reference's registration order (graph_wavenet.py:110-183)."""
    c, d, s, e, k = (cfg.residual_channels, cfg.dilation_channels, cfg.skip_channels,
                     cfg.end_channels, cfg.kernel_size)
    shapes: Dict[str, tuple] = {}
    if cfg.gcn_bool and cfg.adaptive:
        shapes['nodevec1'] = (cfg.num_nodes, 10)
        shapes['nodevec2'] = (10, cfg.num_nodes)
    groups = {
        'filter_convs': [('weight', (d, c, 1, k)), ('bias', (d,))],
        'gate_convs': [('weight', (d, c, 1, k)), ('bias', (d,))],
        'residual_convs': [('weight', (c, d, 1, 1)), ('bias', (c,))],
        'skip_convs': [('weight', (s, d, 1, 1)), ('bias', (s,))],
        'bn': [('weight', (c,)), ('bias', (c,)), ('running_mean', (c,)), ('running_var', (c,)),
               ('num_batches_tracked', ())],
        'gconv': [('mlp.mlp.weight', (c, (cfg.order * cfg.n_supports + 1) * d, 1, 1)),
                  ('mlp.mlp.bias', (c,))],
    }
    if not cfg.gcn_bool:
        del groups['gconv']
    for grp, items in groups.items():
        for i in range(cfg.n_layers):
            for nm, shp in items:
                shapes[f'{grp}.{i}.{nm}'] = shp
    shapes['start_conv.weight'] = (c, cfg.in_dim, 1, 1)
    shapes['start_conv.bias'] = (c,)
    shapes['end_conv_1.weight'] = (e, s, 1, 1)
    shapes['end_conv_1.bias'] = (e,)
    shapes['end_conv_2.weight'] = (cfg.out_dim, e, 1, 1)
    shapes['end_conv_2.bias'] = (cfg.out_dim,)
    return shapes


def synthetic_state_dict(cfg: GWNetConfig, seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic (numpy PCG64) parameters shared by the golden generator and the
    tests, so fixtures need not store weights.  Scales mimic Conv2d's default
    kaiming-uniform magnitude (1/sqrt(fan_in)) so activations stay O(1)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shp in sorted(state_dict_shapes(cfg).items()):
        if name.endswith('num_batches_tracked'):
            sd[name] = torch.zeros((), dtype=torch.int64)
            continue
        if name.endswith('running_mean'):
            arr = 0.1 * rng.standard_normal(shp)
        elif name.endswith('running_var'):
            arr = 1.0 + 0.2 * rng.random(shp)
        elif name.startswith('bn.') and name.endswith('weight'):
            arr = 1.0 + 0.1 * rng.standard_normal(shp)
        elif name.startswith('nodevec'):
            arr = rng.standard_normal(shp)
        elif name.endswith('bias'):
            arr = 0.1 * rng.standard_normal(shp)
        else:
            fan_in = int(np.prod(shp[1:]))
            arr = rng.standard_normal(shp) / np.sqrt(fan_in)
        sd[name] = torch.tensor(np.asarray(arr, dtype=np.float64)).to(dtype)
    return sd


# --------------------------------------------------------------------------- one layer, for per-op parity
def wavenet_layer(res: torch.Tensor, wf, bf, wg, bg, wm, bm, supports: Sequence[torch.Tensor], dilation: int,
                  order: int = 2, dropout_mask: Optional[torch.Tensor] = None, storage=None):
    """One iteration of the reference's layer loop (graph_wavenet.py:220-247) on an already
    batch-normalised input ``res`` [N,C,V,L]: returns (u = gcn(z)+res[..., -L':], z).  ``storage``
    rounds z and every hop (and their gradients) to the 16-bit type as the CUDA data path does."""
    f = dilated_conv(res, wf, bf, dilation)
    g = dilated_conv(res, wg, bg, dilation)
    if storage is None:
        z = torch.tanh(f) * torch.sigmoid(g)
    else:
        z = _sr(_GateSavedRounded.apply(f, g, storage), storage)
    hh = diffusion_conv(z, supports, wm, bm, order, storage)
    if dropout_mask is not None:
        hh = hh * dropout_mask
    return hh + res[:, :, :, -hh.shape[3]:], z
