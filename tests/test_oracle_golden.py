"""The oracle restatement vs golden vectors produced by executing the reference
classes (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle.cases import GOLDEN_CASES, GOLDEN_DIR, case_inputs, case_supports, fl_adjacency
from oracle.graph_oracle import (asym_adj_dense, double_transition, synthetic_directed_graph,
                                 synthetic_knn_graph)
from oracle.gwnet_oracle import (ForwardTrace, adaptive_adjacency, adaptive_adjacency_backward,
                                 dilation_schedule, gate_backward, gwnet_forward, gwnet_forward_literal,
                                 layer_lengths, node_mix, node_mix_backward, receptive_field,
                                 state_dict_shapes, synthetic_state_dict, GWNetConfig)
from conftest import rel_err


def _load(name):
    return np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'), allow_pickle=False)


# ------------------------------------------------------------------ integer / index work: bit exact
def test_asym_adj_bit_exact_vs_reference():
    g = _load('asym_adj')
    fl = fl_adjacency()
    assert np.array_equal(asym_adj_dense(fl.astype(np.float32)), g['fl_f32'])
    assert np.array_equal(asym_adj_dense(fl.astype(np.float64)), g['fl_f64'])
    assert asym_adj_dense(fl.astype(np.float32)).dtype == np.float32
    d = synthetic_directed_graph(67)
    assert np.array_equal(asym_adj_dense(d.astype(np.float32)), g['dir_fwd'])
    assert np.array_equal(asym_adj_dense(d.T.astype(np.float32)), g['dir_bwd'])
    big = synthetic_directed_graph(301, p=0.02, seed=11)
    assert np.array_equal(asym_adj_dense(big.astype(np.float32)), g['big_fwd'])
    with pytest.raises(ValueError):                       # reference raises on int input
        asym_adj_dense(fl)


def test_fl_graph_facts():
    fl = fl_adjacency()
    assert fl.shape == (67, 67) and fl.sum() == 312 and (fl == fl.T).all() and np.trace(fl) == 0
    deg = fl.sum(1)
    assert deg.min() == 2 and deg.max() == 9
    f, b = double_transition(fl)
    assert np.array_equal(f, b)                           # symmetric graph: fwd == bwd
    assert np.allclose(f.sum(1), 1.0, atol=1e-6)


def test_synthetic_graph_deterministic():
    a = synthetic_knn_graph(310)
    b = synthetic_knn_graph(310)
    assert np.array_equal(a, b) and (a == a.T).all() and a.dtype == np.int64 and np.trace(a) == 0
    assert a.sum(1).min() >= 6


def test_schedules():
    c = GWNetConfig(kernel_size=2, blocks=4, layers=2)
    assert receptive_field(c) == 13 and dilation_schedule(c) == [1, 2] * 4
    assert layer_lengths(c, 12)[1:] == [12, 10, 9, 7, 6, 4, 3, 1]
    c5 = GWNetConfig(kernel_size=2, blocks=4, layers=4)
    assert receptive_field(c5) == 61
    assert layer_lengths(c5, 48)[1:] == [60, 58, 54, 46, 45, 43, 39, 31, 30, 28, 24, 16, 15, 13, 9, 1]
    assert receptive_field(GWNetConfig(kernel_size=1)) == 1


@pytest.mark.parametrize('name', list(GOLDEN_CASES))
def test_state_dict_names_match_reference(name):
    g = _load(name)
    cfg = GOLDEN_CASES[name]['cfg']
    assert list(state_dict_shapes(cfg).keys()) == [str(k) for k in g['state_keys']]


# ------------------------------------------------------------------ whole-block parity
@pytest.mark.parametrize('name', ['c1small', 'nopad'])
def test_oracle_aten_path_matches_reference_golden(name, monkeypatch):
    """The ATen-call form of the oracle (what bench.py times as the CPU baseline) is the same function."""
    import oracle.gwnet_oracle as go
    monkeypatch.setattr(go, 'ATEN_PATH', True)
    test_oracle_matches_reference_golden(name)


@pytest.mark.parametrize('name', list(GOLDEN_CASES))
def test_oracle_matches_reference_golden(name):
    c = GOLDEN_CASES[name]
    cfg = c['cfg']
    g = _load(name)
    sd = synthetic_state_dict(cfg, c['seed'])
    for k, v in sd.items():
        if v.is_floating_point() and 'running' not in k:
            v.requires_grad_(True)
    sup = [torch.tensor(s) for s in case_supports(c['supports'])]
    x_np, _ = case_inputs(name)
    x = torch.tensor(x_np, requires_grad=True)
    tr = ForwardTrace()
    if c.get('literal'):
        out = gwnet_forward_literal(sd, x, sup, cfg, c['horizon'], training=True, trace=tr)
    else:
        out = gwnet_forward(sd, x, sup, cfg, training=True, trace=tr)
    assert tuple(out.shape) == g['out_train'].shape
    assert rel_err(out, g['out_train']) < 2e-6
    loss = torch.nn.functional.mse_loss(out, torch.tensor(g['target']))
    assert abs(loss.item() - float(g['loss'])) < 1e-6 * max(1.0, abs(float(g['loss'])))
    loss.backward()
    assert rel_err(x.grad, g['x_grad']) < 2e-5
    wnorm = {}
    for k in sd:
        if f'gradsum/{k}' in g:
            wnorm[k] = float(g[f'gradsum/{k}'][0])
    for k, v in sd.items():
        if f'gradnone/{k}' in g:                      # never used by the reference
            assert v.grad is None or float(v.grad.abs().max()) == 0.0, k
            continue
        if f'gradsum/{k}' not in g:
            continue
        ref_norm = wnorm[k]
        # analytically-zero grads (bias feeding a training-mode BN): absolute tolerance
        # scaled to the same layer's weight-grad norm
        scale = max(ref_norm, wnorm.get(k.replace('bias', 'weight'), 0.0) if k.endswith('bias') else 0.0, 1e-12)
        head = v.grad.detach().double().flatten()[:16].numpy()
        assert np.abs(head - g[f'gradhead/{k}']).max() <= 2e-4 * scale + 1e-9, k
        if f'grad/{k}' in g:
            diff = (v.grad.detach().double() - torch.tensor(g[f'grad/{k}']).double()).norm().item()
            assert diff <= 2e-5 * scale + 1e-9, (k, diff, scale)
    for k in [k for k in g.files if k.startswith('buf/') and 'running' in k]:
        assert rel_err(tr.new_running[k[4:]], g[k]) < 1e-5, k
    # eval mode uses the *updated* running statistics in the golden
    sd_eval = {k: v.detach() for k, v in sd.items()}
    sd_eval.update({k: v for k, v in tr.new_running.items()})
    with torch.no_grad():
        if c.get('literal'):
            oe = gwnet_forward_literal(sd_eval, x.detach(), sup, cfg, c['horizon'], training=False)
        else:
            oe = gwnet_forward(sd_eval, x.detach(), sup, cfg, training=False)
    assert rel_err(oe, g['out_eval']) < 5e-6


# ------------------------------------------------------------------ closed-form backward pieces (fp64)
def test_closed_form_backwards_fp64():
    torch.manual_seed(1)
    e1 = torch.randn(13, 10, dtype=torch.float64, requires_grad=True)
    e2 = torch.randn(10, 13, dtype=torch.float64, requires_grad=True)
    gp = torch.randn(13, 13, dtype=torch.float64)
    adaptive_adjacency(e1, e2).backward(gp)
    d1, d2 = adaptive_adjacency_backward(e1.detach(), e2.detach(), gp)
    assert rel_err(d1, e1.grad) < 1e-12 and rel_err(d2, e2.grad) < 1e-12

    f = torch.randn(5, 7, dtype=torch.float64, requires_grad=True)
    g = torch.randn(5, 7, dtype=torch.float64, requires_grad=True)
    dz = torch.randn(5, 7, dtype=torch.float64)
    (torch.tanh(f) * torch.sigmoid(g)).backward(dz)
    df, dg = gate_backward(f.detach(), g.detach(), dz)
    assert rel_err(df, f.grad) < 1e-12 and rel_err(dg, g.grad) < 1e-12

    x = torch.randn(2, 3, 6, 4, dtype=torch.float64, requires_grad=True)
    a = torch.randn(6, 6, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(2, 3, 6, 4, dtype=torch.float64)
    node_mix(x, a).backward(gy)
    dx, da = node_mix_backward(x.detach(), a.detach(), gy)
    assert rel_err(dx, x.grad) < 1e-12 and rel_err(da, a.grad) < 1e-12
    assert torch.allclose(node_mix(x, a), torch.einsum('ncvl,vw->ncwl', x, a))


def test_oracle_pinned_head_decisions_and_selective_storage():
    """The two checker options the bf16 GPU tests rely on: (1) replacing the head's ReLUs by masks taken from the same
    evaluation changes nothing (output and every gradient identical); (2) `STORAGE_EXCLUDE` removes exactly the named
    rounding points (excluding all of them gives back the exact evaluation)."""
    import oracle.gwnet_oracle as go
    cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=1, layers=2, skip_channels=64,
                      end_channels=64, dropout=0.0)
    sup = [torch.tensor(s).double() for s in case_supports('dir')]
    sd0 = synthetic_state_dict(cfg, 3)
    rng = np.random.default_rng(4)
    x_np = rng.standard_normal((2, 2, 67, 5))

    def run(**kw):
        sd = {k: (v.double().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in sd0.items()}
        tr = go.ForwardTrace()
        out = go.gwnet_forward(sd, torch.tensor(x_np), sup, cfg, training=True, trace=tr, **kw)
        out.square().mean().backward()
        return out.detach(), {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.requires_grad and v.grad is not None}, tr, sd

    out, g, tr, sd = run()
    m1 = (tr.skip.detach() > 0).double()
    e1 = go.pointwise(torch.relu(tr.skip.detach()), sd['end_conv_1.weight'].detach(), sd['end_conv_1.bias'].detach())
    out_p, g_p, _, _ = run(head_masks=(m1, (e1 > 0).double()))
    assert torch.equal(out, out_p)
    for k in g:
        assert torch.allclose(g[k], g_p[k], rtol=1e-12, atol=1e-15), k
    out_s, g_s, _, _ = run(storage=torch.bfloat16)
    assert (out_s - out).norm() / out.norm() > 1e-4                   # rounding is visible ...
    go.STORAGE_EXCLUDE = {'u', 'z', 'hops', 'gate_saved'}
    try:
        out_x, g_x, _, _ = run(storage=torch.bfloat16)
    finally:
        go.STORAGE_EXCLUDE = set()
    assert torch.equal(out_x, out)                                    # ... and gone with every point excluded
    for k in g:
        assert torch.allclose(g[k], g_x[k], rtol=1e-12, atol=1e-15), k
