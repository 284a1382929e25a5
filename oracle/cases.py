"""Golden-case table shared by tests/golden/make_golden.py and the tests.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import os

import numpy as np

from .gwnet_oracle import GWNetConfig
from .graph_oracle import asym_adj_dense, synthetic_directed_graph

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

# name -> (cfg, supports kind, batch, T_in, seed, literal, horizon, full_grads, gcn_bool)
GOLDEN_CASES = {
    # literal reference configuration: B=1, 67 counties, 320 features, k=1, supports=[I]+adp
    'literal': dict(cfg=GWNetConfig(num_nodes=67, in_dim=320, out_dim=256, kernel_size=1,
                                    n_fixed_supports=1, dropout=0.0),
                    supports='eye', n=1, t_in=3, seed=101, literal=True, horizon=3),
    # BASELINE config-1 shape at reduced batch, directed supports (fwd != bwd)
    'c1small': dict(cfg=GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, dropout=0.0),
                    supports='dir', n=4, t_in=12, seed=202),
    # config-5 structure (4x4 layers, dilations to 8, rf=61 > T -> left pad), reduced widths
    'long': dict(cfg=GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=4, layers=4,
                                 skip_channels=64, end_channels=128, dropout=0.0),
                 supports='fl', n=2, t_in=20, seed=303, full_grads=True),
    # T > rf: no padding, L_final = 5; one fixed support; odd widths; k=3
    'nopad': dict(cfg=GWNetConfig(num_nodes=67, in_dim=3, out_dim=5, kernel_size=3, blocks=2, layers=2,
                                  skip_channels=64, end_channels=96, n_fixed_supports=1, dropout=0.0),
                  supports='dir1', n=3, t_in=17, seed=404, full_grads=True),
    # adaptive adjacency off (fixed supports only)
    'noadp': dict(cfg=GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=2, layers=2,
                                  skip_channels=64, end_channels=128, adaptive=False, dropout=0.0),
                  supports='dir', n=2, t_in=12, seed=505, full_grads=True),
    # gcn_bool=False: residual_convs path (graph_wavenet.py:245)
    'nogcn': dict(cfg=GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=2, layers=2,
                                  skip_channels=64, end_channels=128, adaptive=False, n_fixed_supports=0,
                                  gcn_bool=False, dropout=0.0),
                  supports='dir', n=2, t_in=12, seed=606, full_grads=True),
}


def fl_adjacency() -> np.ndarray:
    """The 67x67 Florida county adjacency (fixture copy of the reference's
    data/graph/adj_mx_fl.csv values, written by make_golden.py)."""
    return np.load(os.path.join(GOLDEN_DIR, 'adj_mx_fl.npy')).astype(np.int64)


def case_supports(kind: str, asym=asym_adj_dense):
    """Fixed supports for a golden case, built with ``asym`` (the oracle's
    restatement by default; make_golden passes the reference's own asym_adj)."""
    if kind == 'eye':
        return [np.diag(np.ones(67)).astype(np.float32)]
    if kind == 'fl':
        a = fl_adjacency().astype(np.float32)
        return [np.asarray(asym(a)), np.asarray(asym(a.T.copy()))]
    d = synthetic_directed_graph(67).astype(np.float32)
    sup = [np.asarray(asym(d)), np.asarray(asym(d.T.copy()))]
    return sup[:1] if kind == 'dir1' else sup


def case_inputs(name: str):
    """Seeded input tensor of a golden case (numpy, fp32) — same draw as make_golden."""
    c = GOLDEN_CASES[name]
    cfg = c['cfg']
    rng = np.random.default_rng(c['seed'] + 1)
    if c.get('literal'):
        return rng.standard_normal((cfg.num_nodes, c['horizon'], cfg.in_dim)).astype(np.float32), rng
    return rng.standard_normal((c['n'], cfg.in_dim, cfg.num_nodes, c['t_in'])).astype(np.float32), rng


def synthetic_module_state(state_shapes, seed: int):
    """Deterministic (numpy PCG64) values for an arbitrary module ``state_dict`` given as an ordered list of
    ``(key, shape, dtype_str)``: fan-in scaled weights, BatchNorm scales near 1, positive running variances, zero
    counters.  Shared by tests/golden/make_golden_unet.py (which fills the REFERENCE classes with it) and the tests
    (which fill this repo's modules with it), so no weights need to be stored."""
    import torch
    rng = np.random.default_rng(seed)
    out = {}
    for key, shape, dtype in state_shapes:
        shape = tuple(int(s) for s in shape)
        if key.endswith('num_batches_tracked'):
            out[key] = torch.zeros((), dtype=torch.int64)
            continue
        v = rng.standard_normal(shape if shape else (1,)).astype(np.float32).reshape(shape)
        if key.endswith('running_var'):
            v = np.abs(v) + 0.5
        elif key.endswith('running_mean'):
            v = 0.1 * v
        elif len(shape) == 1 and key.endswith('weight'):          # BatchNorm scale
            v = 1.0 + 0.1 * v
        elif len(shape) == 1:                                      # biases
            v = 0.1 * v
        else:
            fan_in = int(np.prod(shape[1:]))
            v = v / np.sqrt(max(fan_in, 1))
        out[key] = torch.tensor(v)
    return out
