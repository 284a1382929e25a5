"""-m gpu parity tests: the CUDA path (through torch.ops.gwn.* -> libgwn C ABI) against
(1) golden vectors produced by executing the reference classes, and (2) the fp64 CPU oracle on
seeded inputs.  Tolerances are BASELINE.json's: 1e-4 relative (fp32), 2e-2 (bf16), per-tensor
relative L2; bias gradients that are analytically zero (bias feeding a training-mode BN) are
compared with an absolute tolerance scaled to the same layer's weight-gradient norm."""
import os

import numpy as np
import pytest
import torch

from oracle.cases import GOLDEN_CASES, GOLDEN_DIR, case_inputs, case_supports
from oracle.graph_oracle import double_transition, synthetic_directed_graph, synthetic_knn_graph
from oracle.gwnet_oracle import (GWNetConfig, adaptive_adjacency, adaptive_adjacency_backward, layer_lengths,
                                 node_mix)
from gpu_helpers import build_model, captured_head_masks, compare_grads, load_synth, oracle_run, rel

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4
BF16_TOL = 2e-2


def _golden(name):
    return np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'))


@pytest.mark.parametrize('name', list(GOLDEN_CASES))
def test_golden_reference_vectors_fp32(name):
    c = GOLDEN_CASES[name]
    cfg, g = c['cfg'], _golden(name)
    sup = case_supports(c['supports'])
    m = build_model(cfg, sup, horizon=c.get('horizon', 1))
    assert list(m.state_dict().keys()) == [str(k) for k in g['state_keys']]
    sd = load_synth(m, cfg, c['seed'])
    x_np, _ = case_inputs(name)
    x = torch.tensor(x_np, device='cuda', requires_grad=True)
    m.train()
    out = m(x)
    assert tuple(out.shape) == g['out_train'].shape
    assert rel(out, g['out_train']) < FP32_TOL
    loss = torch.nn.functional.mse_loss(out, torch.tensor(g['target'], device='cuda'))
    assert abs(loss.item() - float(g['loss'])) < FP32_TOL * abs(float(g['loss']))
    loss.backward()
    assert rel(x.grad, g['x_grad']) < FP32_TOL
    # gradients: golden keeps norms + first 16 values for all params, full tensors for some cases
    norms = {k[8:]: float(g[k][0]) for k in g.files if k.startswith('gradsum/')}
    for k, p in m.named_parameters():
        if f'gradnone/{k}' in g.files:
            assert p.grad is None, f'{k}: the reference never produces a gradient here'
            continue
        scale = norms[k]
        if k.endswith('bias'):
            scale = max(scale, norms.get(k[:-4] + 'weight', 0.0))
        head = p.grad.detach().double().flatten()[:16].cpu().numpy()
        assert np.abs(head - g[f'gradhead/{k}']).max() <= 4 * FP32_TOL * scale + 1e-12, k
        if f'grad/{k}' in g.files:
            diff = (p.grad.detach().double().cpu() - torch.tensor(g[f'grad/{k}']).double()).norm().item()
            assert diff <= FP32_TOL * scale + 1e-12, (k, diff / scale)
    for k in [k for k in g.files if k.startswith('buf/')]:
        if 'running' in k:
            assert rel(m.state_dict()[k[4:]], g[k]) < 1e-5, k
        else:
            assert int(m.state_dict()[k[4:]]) == int(g[k]), k      # num_batches_tracked: bit exact
    m.eval()
    with torch.no_grad():
        assert rel(m(x.detach()), g['out_eval']) < FP32_TOL


def _oracle_case(cfg, sup, n, t_in, seed, dtype, tol, masks=False, autocast=False):
    m = build_model(cfg, sup)
    sd = load_synth(m, cfg, seed)
    rng = np.random.default_rng(seed + 1)
    x_np = rng.standard_normal((n, cfg.in_dim, cfg.num_nodes, t_in)).astype(np.float32)
    L = layer_lengths(cfg, t_in)
    y_np = rng.standard_normal((n, cfg.out_dim, cfg.num_nodes, L[-1])).astype(np.float32)
    dm_o = dm_g = None
    if masks:
        keep = 1.0 - cfg.dropout
        dm_np = [(rng.random((n, 32, cfg.num_nodes, L[i + 1])) < keep).astype(np.float32) / keep
                 for i in range(cfg.n_layers)]
        dm_o = [torch.tensor(d, dtype=torch.float64) for d in dm_np]
        dm_g = [torch.tensor(d, device='cuda') for d in dm_np]
    bf16 = autocast or dtype == torch.bfloat16
    out_o, loss_o, grads, tr = oracle_run(cfg, sd, x_np, sup, y_np, dropout_masks=dm_o)
    x = torch.tensor(x_np, device='cuda', requires_grad=True)
    m.train()
    from multimodal_outage_b200 import ops
    ops.HEAD_CAPTURE = {} if bf16 else None
    try:
        if autocast:
            with torch.autocast('cuda', dtype=torch.bfloat16):
                out = m(x, dropout_masks=dm_g)
        else:
            m.compute_dtype = dtype
            out = m(x, dropout_masks=dm_g)
    finally:
        cap, ops.HEAD_CAPTURE = ops.HEAD_CAPTURE, None
    assert out.dtype == torch.float32
    assert rel(out, out_o) < tol, rel(out, out_o)
    loss = torch.nn.functional.mse_loss(out, torch.tensor(y_np, device='cuda'))
    assert abs(loss.item() - loss_o.item()) < tol * abs(loss_o.item())
    loss.backward()
    if bf16:
        # Output and loss: 2e-2 against the exact fp64 oracle (above).  Gradients: a ReLU network is piecewise linear in
        # its activations, and ANY 16-bit evaluation flips ~0.15 % of the head's ReLU decisions (measured:
        # profiles/r2_bf16_grad_parity.json), which alone moves per-tensor gradient L2 by 4-10 % at these batch sizes -
        # the bf16-storage oracle shows the same figures, and so does the reference's own autocast run (7.5e-2 .. 1.2e-1,
        # SURVEY App. C.4).  Two FIXED bars:
        #  (1) with the head's ReLU decisions pinned to the ones the CUDA run took (the masks its backward used), every
        #      gradient tensor is within the north-star 2e-2 of the exact fp64 oracle - everything that is smooth is held
        #      to the stated tolerance;
        #  (2) unpinned, every gradient tensor is within 0.12 of the exact oracle (the reference's own bf16-vs-fp32 spread).
        m1, m2 = captured_head_masks(cap, n, cfg.num_nodes, L[-1])
        _, _, grads_p, _ = oracle_run(cfg, sd, x_np, sup, y_np, dropout_masks=dm_o, head_masks=(m1, m2))
        flips = float(((tr.skip.detach() > 0).double() != m1).double().mean())
        worst_p, worst_e, worst_name = 0.0, 0.0, ''
        names = ['__x__'] + [k for k, p in m.named_parameters()
                             if grads.get(k) is not None and not (k.endswith('bias') and 'gconv' in k)]
        for k in names:
            g_k = x.grad if k == '__x__' else dict(m.named_parameters())[k].grad
            e_p, e_e = rel(g_k, grads_p[k]), rel(g_k, grads[k])
            if k.endswith('bias') and k != '__x__' and not k.startswith('bn.'):
                # conv biases feeding a training-mode BatchNorm have an analytically zero gradient (noise / noise)
                wk = k[:-4] + 'weight'
                scale = float(grads_p[wk].norm()) if grads_p.get(wk) is not None else 0.0
                if float(grads_p[k].norm()) < 1e-6 * max(scale, 1e-30):
                    continue
            if e_p > worst_p:
                worst_p, worst_name = e_p, k
            worst_e = max(worst_e, e_e)
        print(f'bf16 whole-model: out {rel(out, out_o):.2e}; {100 * flips:.2f}% head ReLU decisions flipped; worst gradient '
              f'tensor vs exact fp64: {worst_e:.2e} unpinned, {worst_p:.2e} with the head decisions pinned ({worst_name})')
        assert worst_p < BF16_TOL, (worst_name, worst_p)
        assert worst_e < 0.12, worst_e
        return m, []
    assert rel(x.grad, grads.pop('__x__')) < tol
    rep = []
    bad = compare_grads(m, grads, tol, rep)
    assert not bad, bad
    for k, v in tr.new_running.items():
        assert rel(m.state_dict()[k], v) < max(tol, 1e-5), k
    return m, rep


def test_fp32_vs_fp64_oracle_directed_c1_shape():
    cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, dropout=0.0)
    sup = case_supports('dir')
    _oracle_case(cfg, sup, n=8, t_in=12, seed=11, dtype=torch.float32, tol=FP32_TOL)


def test_bf16_vs_fp64_oracle_c2_shape():
    """bf16 activation storage (BASELINE config 2 at reduced batch) within 2e-2 of the fp64 oracle."""
    cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, dropout=0.0)
    sup = double_transition(np.load(os.path.join(GOLDEN_DIR, 'adj_mx_fl.npy')).astype(np.float32))
    _oracle_case(cfg, sup, n=32, t_in=12, seed=12, dtype=torch.bfloat16, tol=BF16_TOL)


def test_bf16_via_autocast():
    cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=2, layers=2, dropout=0.0)
    _oracle_case(cfg, case_supports('dir'), n=16, t_in=12, seed=13, dtype=None, tol=BF16_TOL, autocast=True)


def test_explicit_dropout_masks_match_oracle():
    cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=2, layers=2, dropout=0.3)
    _oracle_case(cfg, case_supports('dir'), n=4, t_in=12, seed=14, dtype=torch.float32, tol=FP32_TOL, masks=True)


def test_larger_graph_310_nodes_fp32():
    """Generic-V kernels (tiles of 64 nodes, ragged edge) on a seeded 310-node kNN graph."""
    cfg = GWNetConfig(num_nodes=310, in_dim=2, out_dim=12, kernel_size=2, blocks=2, layers=2,
                      skip_channels=64, end_channels=128, dropout=0.0)
    sup = double_transition(synthetic_knn_graph(310))
    _oracle_case(cfg, sup, n=2, t_in=12, seed=15, dtype=torch.float32, tol=FP32_TOL)


def test_wide_input_in_dim_320_and_long_T():
    cfg = GWNetConfig(num_nodes=67, in_dim=320, out_dim=256, kernel_size=2, blocks=2, layers=2,
                      skip_channels=64, end_channels=64, dropout=0.0)
    _oracle_case(cfg, case_supports('fl'), n=3, t_in=9, seed=16, dtype=torch.float32, tol=FP32_TOL)


def test_wide_input_in_dim_320_bf16_start_conv_on_tensor_cores():
    """Config-4 style input (256 features + 64 date2vec channels): the start conv runs as a TMA-fed tcgen05 GEMM on
    a channels-last bf16 copy of the input; whole-model bf16 bar incl. the gradient wrt the input."""
    cfg = GWNetConfig(num_nodes=67, in_dim=320, out_dim=256, kernel_size=2, blocks=2, layers=2,
                      skip_channels=64, end_channels=64, dropout=0.0)
    _oracle_case(cfg, case_supports('fl'), n=3, t_in=9, seed=16, dtype=torch.bfloat16, tol=BF16_TOL)
    _oracle_case(cfg, case_supports('fl'), n=2, t_in=3, seed=17, dtype=torch.bfloat16, tol=BF16_TOL)    # padded time


def test_fused_dropout_is_a_valid_mask():
    cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=1, layers=2, dropout=0.3)
    m = build_model(cfg, case_supports('dir'))
    load_synth(m, cfg, 17)
    x = torch.randn(16, 2, 67, 12, device='cuda')
    m.train()
    torch.manual_seed(5)
    m._rng_state = None
    o1 = m(x)
    m._rng_state = None
    o2 = m(x)
    assert torch.equal(o1, o2)                     # same {seed, offset} -> same mask
    o3 = m(x)
    assert not torch.equal(o1, o3)                 # offset advanced -> fresh mask
    m.dropout = 0.0
    o0 = m(x)
    assert rel(o1, o0) > 1e-3
    # gradient flows through the same mask as forward (finite difference on one weight)
    m.dropout = 0.3
    w = m.gconv[0].mlp.mlp.weight
    def f():
        m._rng_state = None
        return m(x).double().pow(2).sum()
    m.zero_grad(); f().backward()
    g = w.grad[3, 40, 0, 0].item()
    eps = 1e-2
    with torch.no_grad():
        w[3, 40, 0, 0] += eps; fp = f().item(); w[3, 40, 0, 0] -= 2 * eps; fm = f().item(); w[3, 40, 0, 0] += eps
    assert abs((fp - fm) / (2 * eps) - g) < 5e-2 * max(1.0, abs(g))


def test_fused_dropout_same_mask_in_every_kernel_path():
    """The in-kernel Philox mask is a function of (seed, offset, element) only: the fp32 CUDA-core path, the
    unfused bf16 tensor-core path and the fused bf16 gcn kernel must all draw the SAME mask, forward and
    backward (so bf16 output/gradients stay within the bf16 bar of the fp32 run with dropout on)."""
    cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=1, layers=2, dropout=0.3)
    m = build_model(cfg, case_supports('dir'))
    load_synth(m, cfg, 19)
    x = torch.randn(8, 2, 67, 12, device='cuda')
    m.train()
    outs, grads = [], []
    for dt in (torch.float32, torch.bfloat16):
        m.compute_dtype = dt
        m._rng_state = None
        torch.manual_seed(5)
        m.zero_grad()
        o = m(x)
        o.square().mean().backward()
        outs.append(o.detach().clone())
        grads.append(m.gconv[0].mlp.mlp.weight.grad.detach().clone())
    assert rel(outs[1], outs[0]) < BF16_TOL, rel(outs[1], outs[0])
    assert rel(grads[1], grads[0]) < 0.1, rel(grads[1], grads[0])


def test_adaptive_adjacency_op_fwd_bwd():
    from multimodal_outage_b200 import ops
    for V in (67, 310, 5):
        torch.manual_seed(V)
        e1 = torch.randn(V, 10, dtype=torch.float64)
        e2 = torch.randn(10, V, dtype=torch.float64)
        gp = torch.randn(V, V, dtype=torch.float64)
        p_ref = adaptive_adjacency(e1, e2)
        d1, d2 = adaptive_adjacency_backward(e1, e2, gp)
        a = e1.float().cuda().requires_grad_(True)
        b = e2.float().cuda().requires_grad_(True)
        p = ops.AdaptiveAdjacency.apply(a, b)
        assert rel(p, p_ref) < 1e-5
        assert torch.allclose(p.sum(1), torch.ones(V, device='cuda'), atol=1e-5)
        p.backward(gp.float().cuda())
        assert rel(a.grad, d1) < FP32_TOL and rel(b.grad, d2) < FP32_TOL


def test_node_mix_primitive_and_nconv_module():
    from multimodal_outage_b200 import nconv, ops
    torch.manual_seed(3)
    for V in (67, 130):
        x = torch.randn(2, 32, V, 5, dtype=torch.float64)
        A = torch.rand(V, V, dtype=torch.float64)
        ref = node_mix(x, A)
        y = nconv()(x.float().cuda(), A.float().cuda())
        assert y.shape == ref.shape and y.is_contiguous()
        assert rel(y, ref) < 1e-5
        xs = x.permute(0, 3, 2, 1).reshape(-1, V, 32).float().cuda().contiguous()
        yt = ops.node_mix(xs, A.float().cuda(), True)
        ref_t = torch.einsum('svc,wv->swc', xs.double().cpu(), A)
        assert rel(yt, ref_t) < 1e-5
        # linearity (size-independent property)
        y2 = ops.node_mix(2.5 * xs, A.float().cuda(), True)
        assert rel(y2, 2.5 * yt) < 1e-6


def test_eval_mode_is_batch_independent_at_full_config2_size():
    """BASELINE config 2 shape (N=512, V=67, T=12, bf16): in eval mode every sample is independent, so a
    slice of the batch must reproduce the same rows (size-independent property at full size)."""
    cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, dropout=0.3)
    m = build_model(cfg, case_supports('fl'))
    load_synth(m, cfg, 18)
    m.compute_dtype = torch.bfloat16
    m.eval()
    x = torch.randn(512, 2, 67, 12, device='cuda')
    with torch.no_grad():
        full = m(x)
        part = m(x[100:164].contiguous())
    assert full.shape == (512, 12, 67, 1)
    assert torch.equal(full[100:164], part)
    assert torch.isfinite(full).all()


def test_training_step_contract_and_adam_step():
    """lit.py:29-43,59-61 contract: training_step(batch) -> scalar loss; Adam(lr=1e-3) step changes weights
    and reduces the loss on a fixed batch."""
    from multimodal_outage_b200.lit import LitGWNet
    cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=2, layers=2, dropout=0.0)
    m = build_model(cfg, case_supports('fl'))
    lit = LitGWNet(m)
    opt = lit.configure_optimizers()['optimizer']
    torch.manual_seed(0)
    x = torch.randn(8, 2, 67, 12, device='cuda'); y = torch.randn(8, 12, 67, 6, device='cuda')
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = lit.training_step((x, y))
        assert loss.dim() == 0
        loss.backward(); opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    assert 'train_loss' in lit.logged and 'train_mae' in lit.logged


def test_non_b200_or_cpu_input_raises():
    from multimodal_outage_b200._lib import GwnError
    cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=1, layers=1, dropout=0.0)
    m = build_model(cfg, case_supports('fl'))
    with pytest.raises(GwnError):
        m(torch.randn(1, 2, 67, 12))


def _layer_op_case(dtype, tol, V=67, N=6, Lin=9, dil=2, taps=2, n_sup=3, with_bn=True, mask=False, seed=0,
                   tensor_cores=False, adp_grad=True, sparse=False):
    """ops.WaveNetLayer (bn-fold + gate + hops + mlp + dropout + residual, fwd AND bwd) against the
    oracle's single-layer restatement on IDENTICAL inputs (no ReLU anywhere -> bf16 error stays at
    rounding level, so the 2e-2 / 1e-4 bars apply to every gradient)."""
    from multimodal_outage_b200 import ops
    from oracle.gwnet_oracle import wavenet_layer, batch_norm_train
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)   # noqa: E731
    Lout = Lin - dil * (taps - 1)
    Lf = min(2, Lout)
    u_prev = rn(N, 32, V, Lin).to(dtype).double()                      # stored (rounded) pre-BN stream
    gamma, beta = 1 + 0.1 * rn(32), 0.1 * rn(32)
    wf, wg = rn(32, 32, 1, taps) / 8, rn(32, 32, 1, taps) / 8
    bf, bg = 0.1 * rn(32), 0.1 * rn(32)
    mlp_in = 32 * (1 + 2 * n_sup)
    wm, bm = rn(32, mlp_in, 1, 1) / mlp_in ** 0.5, 0.1 * rn(32)
    sups = [torch.softmax(rn(V, V), dim=1) for _ in range(n_sup)]
    if sparse:      # fixed supports as the configurations have them: transition matrices of a kNN graph (a few non-zeros per row)
        tr = double_transition(synthetic_knn_graph(V))
        for i in range(min(n_sup - 1, 2)):
            sups[i] = torch.tensor(np.asarray(tr[i]), dtype=torch.float64)
    du = rn(N, 32, V, Lout).to(dtype).double()
    dzl = rn(N, 32, V, Lf).to(dtype).double()
    keep = 0.7
    dm = ((torch.rand(N, 32, V, Lout, generator=g) < keep).double() / keep) if mask else None
    storage = None if dtype == torch.float32 else dtype
    # ---- oracle (fp64, autograd)
    leaves = [t.clone().requires_grad_(True) for t in (u_prev, gamma, beta, wf, bf, wg, bg, wm, bm, sups[-1])]
    o_u, o_g, o_b, o_wf, o_bf, o_wg, o_bg, o_wm, o_bm, o_adp = leaves
    if with_bn:
        res, mean, var = batch_norm_train(o_u, o_g, o_b, 1e-5)
    else:
        res = o_u
    u_o, z_o = wavenet_layer(res, o_wf, o_bf, o_wg, o_bg, o_wm, o_bm, sups[:-1] + [o_adp], dil, 2, dm, storage)
    ((u_o * du).sum() + (z_o[..., -Lf:] * dzl).sum()).backward()
    # ---- CUDA
    cl = lambda t: t.permute(0, 3, 2, 1).contiguous().to(dtype).cuda()   # noqa: E731  NCHW -> [N,L,V,C]
    f32 = lambda t: t.float().cuda()                                     # noqa: E731
    w_fg = torch.stack([wf[:, :, 0, :], wg[:, :, 0, :]], dim=1).permute(3, 2, 0, 1).reshape(taps * 32, 64)
    b_fg = torch.stack([bf, bg], dim=1).reshape(64)
    k_in = [cl(u_prev).requires_grad_(True)]
    if with_bn:
        cnt = N * V * Lin
        stats = torch.stack([u_prev.sum(dim=(0, 2, 3)), (u_prev ** 2).sum(dim=(0, 2, 3))]).cuda()
        gk, bk = f32(gamma).requires_grad_(True), f32(beta).requires_grad_(True)
        rm, rv = torch.zeros(32, device='cuda'), torch.ones(32, device='cuda')
    else:
        stats = gk = bk = rm = rv = None
    wfg_k, bfg_k = f32(w_fg).contiguous().requires_grad_(True), f32(b_fg).requires_grad_(True)
    wm_k, bm_k = f32(wm[:, :, 0, 0].t()).contiguous().requires_grad_(True), f32(bm).requires_grad_(True)
    sup_k = [f32(a) for a in sups]
    if sparse:
        assert all(ops.register_sparse_support(a) for a in sup_k[:min(n_sup - 1, 2)])
        assert not ops.register_sparse_support(sup_k[-1])          # the dense (softmax) support stays a GEMM
    sup_k[-1].requires_grad_(adp_grad)
    meta = dict(training=True, momentum=0.1, eps=1e-5, Lf=Lf, taps=taps, dilation=dil, order=2, has_gconv=True,
                dropout_p=0.3 if mask else 0.0, seed=0, offset=0)
    hop_mats_k = None
    if tensor_cores:      # V <= 80: supports resident on chip; larger: TMA-tiled GEMM per hop
        hop_mats_k = (ops.hop_mats if ops.hop_mode(V, len(sup_k)) == 1 else ops.support_images)([a.detach() for a in sup_k])
    u_k, stats_k, zl_k = ops.WaveNetLayer.apply(k_in[0], stats, gk, bk, rm, rv, wfg_k, bfg_k, wm_k, bm_k,
                                                cl(dm) if mask else None, None, hop_mats_k, meta, *sup_k)
    assert rel(u_k.permute(0, 3, 2, 1), u_o) < tol and rel(zl_k.permute(0, 3, 2, 1), z_o[..., -Lf:]) < tol
    cnt_o = N * V * Lout
    assert rel(stats_k[0] / cnt_o, u_o.mean(dim=(0, 2, 3))) < max(tol, 1e-5) * 10
    torch.autograd.backward([u_k, zl_k], [cl(du), cl(dzl)])
    errs = {
        'du_prev': rel(k_in[0].grad.permute(0, 3, 2, 1), o_u.grad),
        'dw_filter': rel(wfg_k.grad.reshape(taps, 32, 32, 2)[..., 0].permute(2, 1, 0), o_wf.grad[:, :, 0, :]),
        'dw_gate': rel(wfg_k.grad.reshape(taps, 32, 32, 2)[..., 1].permute(2, 1, 0), o_wg.grad[:, :, 0, :]),
        'db_filter': rel(bfg_k.grad.reshape(32, 2)[:, 0], o_bf.grad),
        'db_gate': rel(bfg_k.grad.reshape(32, 2)[:, 1], o_bg.grad),
        'dw_mlp': rel(wm_k.grad.t(), o_wm.grad[:, :, 0, 0]),
        'db_mlp': rel(bm_k.grad, o_bm.grad),
    }
    if adp_grad:
        errs['d_adp'] = rel(sup_k[-1].grad, o_adp.grad)
    if with_bn:
        errs['dgamma'] = rel(gk.grad, o_g.grad)
        errs['dbeta'] = rel(bk.grad, o_b.grad)
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, (bad, errs)
    return errs


@pytest.mark.parametrize('with_bn,mask', [(True, False), (False, True)])
def test_fp32_layer_op_fwd_bwd_vs_oracle(with_bn, mask):
    _layer_op_case(torch.float32, FP32_TOL, with_bn=with_bn, mask=mask, seed=1)
    _layer_op_case(torch.float32, FP32_TOL, V=130, N=2, Lin=5, dil=1, taps=3, n_sup=2, with_bn=with_bn, mask=mask, seed=2)


@pytest.mark.parametrize('tensor_cores', [False, True])
@pytest.mark.parametrize('with_bn,mask', [(True, False), (False, True)])
def test_bf16_layer_op_fwd_bwd_vs_oracle(with_bn, mask, tensor_cores):
    """bf16 storage: every output and EVERY gradient of the layer within 2e-2 of the fp64 oracle, with the
    diffusion hops on CUDA cores and on the tcgen05 tensor cores."""
    errs = _layer_op_case(torch.bfloat16, BF16_TOL, with_bn=with_bn, mask=mask, seed=3, tensor_cores=tensor_cores)
    print(f'bf16 layer-op errors (tensor_cores={tensor_cores}):', {k: f'{v:.1e}' for k, v in errs.items()})
    if tensor_cores:      # ragged tile (slabs not a multiple of 8), 2 supports, 1 support
        _layer_op_case(torch.bfloat16, BF16_TOL, V=67, N=3, Lin=7, dil=1, taps=2, n_sup=2, with_bn=with_bn,
                       mask=mask, seed=4, tensor_cores=True)
        _layer_op_case(torch.bfloat16, BF16_TOL, V=40, N=5, Lin=4, dil=1, taps=2, n_sup=1, with_bn=with_bn,
                       mask=mask, seed=5, tensor_cores=True)
        # kernel size 3: run-time chunk counts in the position GEMMs, separate (unfused) gate weight-gradient launch
        _layer_op_case(torch.bfloat16, BF16_TOL, V=67, N=3, Lin=9, dil=2, taps=3, n_sup=3, with_bn=with_bn,
                       mask=mask, seed=6, tensor_cores=True)


@pytest.mark.parametrize('with_bn,mask', [(True, False), (False, True)])
def test_bf16_layer_op_big_graph_tma_gemm_hops(with_bn, mask):
    """Graphs whose supports do not fit on chip (the 3,100-node configs' code path, at sizes the fp64
    oracle finishes quickly): hops, their backward and dA on the TMA-tiled tcgen05 GEMM; ragged node tiles
    (V % 128 != 0, V % 64 != 0, V % 8 != 0) and odd slab counts."""
    errs = _layer_op_case(torch.bfloat16, BF16_TOL, V=150, N=3, Lin=5, dil=1, taps=2, n_sup=3, with_bn=with_bn,
                          mask=mask, seed=6, tensor_cores=True)
    print('bf16 big-graph layer-op errors V=150:', {k: f'{v:.1e}' for k, v in errs.items()})
    errs = _layer_op_case(torch.bfloat16, BF16_TOL, V=333, N=1, Lin=4, dil=1, taps=2, n_sup=2, with_bn=with_bn,
                          mask=mask, seed=7, tensor_cores=True)
    print('bf16 big-graph layer-op errors V=333:', {k: f'{v:.1e}' for k, v in errs.items()})


@pytest.mark.parametrize('with_bn,mask', [(True, False), (False, True)])
def test_bf16_layer_op_fused_diffusion_backward(with_bn, mask):
    """Supports without gradient (fixed transition matrices only): the diffusion backward runs as ONE fused
    kernel (gcn_fused_bwd.cu: mask, transposed hops, mlp data + weight gradients, gate backward on chip).  Every
    output and gradient within the bf16 bar; ragged slab counts, 1-3 supports, V below and above 64."""
    for kw in (dict(seed=21), dict(V=67, N=3, Lin=7, dil=1, n_sup=2, seed=22), dict(V=40, N=5, Lin=4, dil=1, n_sup=1, seed=23),
               dict(V=80, N=2, Lin=6, dil=2, n_sup=3, seed=24)):
        errs = _layer_op_case(torch.bfloat16, BF16_TOL, taps=2, with_bn=with_bn, mask=mask, tensor_cores=True,
                              adp_grad=False, **kw)
        print('bf16 fused-backward layer-op errors:', kw, {k: f'{v:.1e}' for k, v in errs.items()})


def test_tma_gemm_every_staging_mode():
    """gwn_gemm_test: C = A.B^T through K-major/MN-major 128B-swizzled and slab (64B-swizzled) operands,
    ragged M/N/K, split-K with atomic accumulation."""
    from multimodal_outage_b200 import _lib
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device='cuda').manual_seed(3)
    pad8 = lambda n: (n + 7) // 8 * 8    # noqa: E731
    for (M, N, K, bn, splits) in ((128, 256, 64, 256, 1), (300, 320, 200, 64, 1), (257, 512, 1000, 256, 3)):
        A = torch.randn(M, K, device='cuda', generator=g).to(torch.bfloat16)
        B = torch.randn(N, K, device='cuda', generator=g).to(torch.bfloat16)
        ref = A.float() @ B.float().t()
        for am in (0, 1):
            for bm in (0, 1, 2):
                if am == 0:
                    lda = pad8(K); Ag = torch.zeros(M, lda, device='cuda', dtype=torch.bfloat16); Ag[:, :K] = A
                else:
                    lda = pad8(M); Ag = torch.zeros(K, lda, device='cuda', dtype=torch.bfloat16); Ag[:, :M] = A.t()
                if bm == 0:
                    ldb = pad8(K); Bg = torch.zeros(N, ldb, device='cuda', dtype=torch.bfloat16); Bg[:, :K] = B
                elif bm == 1:
                    ldb = pad8(N); Bg = torch.zeros(K, ldb, device='cuda', dtype=torch.bfloat16); Bg[:, :N] = B.t()
                else:
                    ldb = 32; Bg = B.reshape(N // 32, 32, K).permute(0, 2, 1).contiguous()
                Cc = torch.zeros(M, N, device='cuda')
                _lib.check(lib.gwn_gemm_test(Ag.data_ptr(), Bg.data_ptr(), Cc.data_ptr(), M, N, K, am, bm, lda, ldb, bn,
                                             splits, st), 'gwn_gemm_test')
                assert rel(Cc, ref) < 1e-5, (M, N, K, am, bm, rel(Cc, ref))


def test_big_graph_hop_and_support_gradient_kernels():
    """gwn_hop_big (both operand images, with and without add-in) and gwn_dadj_big against fp64 einsums."""
    import ctypes as C
    from multimodal_outage_b200 import ops, _lib
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(1)
    for V, slabs in ((150, 5), (333, 16), (1000, 9)):
        sups = [torch.softmax(torch.randn(V, V, device='cuda'), dim=1) for _ in range(2)]
        img = ops.support_images(sups)
        x = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
        add = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
        y = torch.empty_like(x)
        for s in range(2):
            Ab = sups[s].to(torch.bfloat16).double()
            for which in range(2):
                ref = torch.einsum('vw,svc->swc' if which == 0 else 'wv,svc->swc', Ab, x.double())
                _lib.check(lib.gwn_hop_big(img.data_ptr(), 2, s, which, x.data_ptr(), y.data_ptr(), None, slabs, V, st), 'hop')
                assert rel(y, ref) < 4e-3, (V, slabs, s, which, rel(y, ref))
                _lib.check(lib.gwn_hop_big(img.data_ptr(), 2, s, which, x.data_ptr(), y.data_ptr(), add.data_ptr(), slabs,
                                           V, st), 'hop')
                assert rel(y, ref + add.double()) < 4e-3
        dA = torch.ones(V, V, device='cuda')
        _lib.check(lib.gwn_dadj_big(x.data_ptr(), add.data_ptr(), dA.data_ptr(), slabs, V, st), 'dadj')
        ref = 1.0 + torch.einsum('svc,swc->vw', x.double(), add.double())
        assert rel(dA, ref) < 1e-5, (V, slabs, rel(dA, ref))


def test_bf16_whole_model_big_graph_310_nodes():
    """Whole gwnet in bf16 on a 310-node kNN graph: the TMA-GEMM hop path end to end (fwd + bwd)."""
    cfg = GWNetConfig(num_nodes=310, in_dim=2, out_dim=12, kernel_size=2, blocks=2, layers=2,
                      skip_channels=64, end_channels=128, dropout=0.0)
    sup = double_transition(synthetic_knn_graph(310))
    _oracle_case(cfg, sup, n=2, t_in=12, seed=16, dtype=torch.bfloat16, tol=BF16_TOL)


def test_bf16_whole_model_config3_structure_at_3100_nodes():
    """BASELINE config 3's model (4 x 2 layers, fwd / bwd transition matrices of a 3,100-node kNN graph + adaptive support,
    dropout masks) at the REAL graph size, forward + backward against the fp64 oracle: sparse fixed-support hops, the
    two-SM hop GEMM (layers with >= 8 slabs) and the single-SM one (the short layers), Horner backward, support gradient.
    Two samples keep the oracle's 3100 x 3100 products in the seconds."""
    cfg = GWNetConfig(num_nodes=3100, in_dim=2, out_dim=12, kernel_size=2, blocks=4, layers=2, dropout=0.3)
    sup = double_transition(synthetic_knn_graph(3100))
    _oracle_case(cfg, sup, n=2, t_in=12, seed=17, dtype=torch.bfloat16, tol=BF16_TOL, masks=True)


def test_tc_hop_kernel_every_image_variant():
    """gwn_hop_tc (tcgen05) against an fp64 einsum for A^T, (A^2)^T, A, A^2 images, ragged slab counts."""
    from multimodal_outage_b200 import ops
    torch.manual_seed(0)
    for V, slabs in ((67, 8), (67, 37), (80, 5), (33, 300)):
        sups = [torch.softmax(torch.randn(V, V, device='cuda'), dim=1) for _ in range(2)]
        mats = ops.hop_mats(sups)
        buf = torch.randn(slabs * V, 96, device='cuda').to(torch.bfloat16)
        x = buf[:, :32].double().reshape(slabs, V, 32).clone()
        for s in range(2):
            A = sups[s].double()
            refs = [torch.einsum('vw,svc->swc', A, x), torch.einsum('vw,svc->swc', A @ A, x),
                    torch.einsum('wv,svc->swc', A, x), torch.einsum('wv,svc->swc', A @ A, x)]
            for variant in range(4):
                ops.hop_tc(mats, 8, 4 * s + variant, buf, 0, 1 + variant % 2, V)
                y = buf[:, 32 * (1 + variant % 2):32 * (2 + variant % 2)].double().reshape(slabs, V, 32)
                assert rel(y, refs[variant]) < 1e-2, (V, slabs, s, variant)
        assert torch.equal(buf[:, :32].double().reshape(slabs, V, 32), x)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_skip_head_op_fwd_bwd_vs_oracle(dtype):
    """relu(sum_i Ws_i z_i + bs) -> relu(end1) -> end2 on identical (already stored) z_last inputs: the head
    computes in fp32 in both modes, so it must meet the fp32 bar even when z_last is bf16."""
    from multimodal_outage_b200 import ops
    from oracle.gwnet_oracle import pointwise
    g = torch.Generator().manual_seed(5)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)   # noqa: E731
    N, V, Lf, nl, S, E, O = 5, 67, 2, 3, 64, 96, 12
    zs = [torch.tanh(rn(N, 32, V, Lf)).to(dtype).double().requires_grad_(True) for _ in range(nl)]
    Ws = [(rn(S, 32, 1, 1) / 6).requires_grad_(True) for _ in range(nl)]
    bs = [(0.1 * rn(S)).requires_grad_(True) for _ in range(nl)]
    W1, b1 = (rn(E, S, 1, 1) / 8).requires_grad_(True), (0.1 * rn(E)).requires_grad_(True)
    W2, b2 = (rn(O, E, 1, 1) / 10).requires_grad_(True), (0.1 * rn(O)).requires_grad_(True)
    skip = sum(pointwise(z, w, b) for z, w, b in zip(zs, Ws, bs))
    out_o = pointwise(torch.relu(pointwise(torch.relu(skip), W1, b1)), W2, b2)
    dout = rn(N, O, V, Lf)
    out_o.backward(dout)
    f32 = lambda t: t.detach().float().cuda()   # noqa: E731
    w_skip = f32(torch.cat([w[:, :, 0, 0].t() for w in Ws], 0)).contiguous().requires_grad_(True)
    b_skip = f32(sum(bs)).requires_grad_(True)
    w_end1, b_end1 = f32(W1[:, :, 0, 0].t()).contiguous().requires_grad_(True), f32(b1).requires_grad_(True)
    w2p = torch.zeros(E, 32); w2p[:, :O] = W2.detach()[:, :, 0, 0].t().float()
    b2p = torch.zeros(32); b2p[:O] = b2.detach().float()
    w_end2, b_end2 = w2p.cuda().requires_grad_(True), b2p.cuda().requires_grad_(True)
    zk = [z.detach().permute(0, 3, 2, 1).contiguous().to(dtype).cuda().requires_grad_(True) for z in zs]
    out_k = ops.SkipHead.apply(w_skip, b_skip, w_end1, b_end1, w_end2, b_end2, O, *zk)
    assert rel(out_k, out_o) < FP32_TOL
    out_k.backward(dout.float().cuda())
    gtol = FP32_TOL if dtype == torch.float32 else BF16_TOL      # dz_last is stored in `dtype`
    for i in range(nl):
        assert rel(zk[i].grad.permute(0, 3, 2, 1), zs[i].grad) < gtol
        assert rel(w_skip.grad[32 * i:32 * i + 32].t(), Ws[i].grad[:, :, 0, 0]) < FP32_TOL
    assert rel(b_skip.grad, bs[0].grad) < FP32_TOL
    assert rel(w_end1.grad.t(), W1.grad[:, :, 0, 0]) < FP32_TOL and rel(b_end1.grad, b1.grad) < FP32_TOL
    assert rel(w_end2.grad[:, :O].t(), W2.grad[:, :, 0, 0]) < FP32_TOL and rel(b_end2.grad[:O], b2.grad) < FP32_TOL


@pytest.mark.parametrize('nl,S,E,Lf', [(3, 128, 256, 2), (8, 256, 512, 1)])
def test_bf16_head_on_tensor_cores_fwd_bwd_vs_oracle(nl, S, E, Lf):
    """bf16 head as TMA-fed tcgen05 GEMMs (csrc/head_tc.cu): bf16 operands, fp32 accumulation; every output and
    gradient within the bf16 bar of the fp64 oracle, with arbitrary fp32 weights (the GEMMs that feed a ReLU run
    in split hi/lo bf16 precision so relu masks do not flip)."""
    from multimodal_outage_b200 import ops
    from oracle.gwnet_oracle import pointwise
    g = torch.Generator().manual_seed(9)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)   # noqa: E731
    r16 = lambda t: t.to(torch.bfloat16).double()                       # noqa: E731
    N, V, O = 7, 67, 12
    zs = [r16(torch.tanh(rn(N, 32, V, Lf))).requires_grad_(True) for _ in range(nl)]
    f32r = lambda t: t.float().double()                                 # noqa: E731  (parameters are fp32)
    Ws = [f32r(rn(S, 32, 1, 1) / 6).requires_grad_(True) for _ in range(nl)]
    bs = [(0.1 * rn(S)).requires_grad_(True) for _ in range(nl)]
    W1, b1 = f32r(rn(E, S, 1, 1) / S ** 0.5).requires_grad_(True), (0.1 * rn(E)).requires_grad_(True)
    W2, b2 = f32r(rn(O, E, 1, 1) / E ** 0.5).requires_grad_(True), (0.1 * rn(O)).requires_grad_(True)
    skip = sum(pointwise(z, w, b) for z, w, b in zip(zs, Ws, bs))
    out_o = pointwise(torch.relu(pointwise(torch.relu(skip), W1, b1)), W2, b2)
    dout = rn(N, O, V, Lf)
    out_o.backward(dout)
    f32 = lambda t: t.detach().float().cuda()   # noqa: E731
    w_skip = f32(torch.cat([w[:, :, 0, 0].t() for w in Ws], 0)).contiguous().requires_grad_(True)
    b_skip = f32(sum(bs)).requires_grad_(True)
    w_end1, b_end1 = f32(W1[:, :, 0, 0].t()).contiguous().requires_grad_(True), f32(b1).requires_grad_(True)
    w2p = torch.zeros(E, 32); w2p[:, :O] = W2.detach()[:, :, 0, 0].t().float()
    b2p = torch.zeros(32); b2p[:O] = b2.detach().float()
    w_end2, b_end2 = w2p.cuda().requires_grad_(True), b2p.cuda().requires_grad_(True)
    zk = [z.detach().permute(0, 3, 2, 1).contiguous().to(torch.bfloat16).cuda().requires_grad_(True) for z in zs]
    assert ops.head_tc_supported(N * V * Lf, S, E)
    out_k = ops.SkipHead.apply(w_skip, b_skip, w_end1, b_end1, w_end2, b_end2, O, *zk)
    errs = {'out': rel(out_k, out_o)}
    out_k.backward(dout.float().cuda())
    for i in range(nl):
        errs[f'dz{i}'] = rel(zk[i].grad.permute(0, 3, 2, 1), zs[i].grad)
        errs[f'dWs{i}'] = rel(w_skip.grad[32 * i:32 * i + 32].t(), Ws[i].grad[:, :, 0, 0])
    errs['dbs'] = rel(b_skip.grad, bs[0].grad)
    errs['dW1'] = rel(w_end1.grad.t(), W1.grad[:, :, 0, 0]); errs['db1'] = rel(b_end1.grad, b1.grad)
    errs['dW2'] = rel(w_end2.grad[:, :O].t(), W2.grad[:, :, 0, 0]); errs['db2'] = rel(b_end2.grad[:O], b2.grad)
    print('bf16 tensor-core head errors:', {k: f'{v:.1e}' for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v < BF16_TOL}
    assert not bad, bad


@pytest.mark.gpu
def test_param_packing_layouts_and_gradient_unpacking():
    """Parameter packing = one gather launch; gradient scatter = one launch per autograd node (_PackLayers, _PackHead;
    csrc/pack.cu): check layouts element-wise and that gradients of the packed tensors are routed back to the
    reference-shaped parameters exactly."""
    from multimodal_outage_b200 import gwnet
    torch.manual_seed(0)
    m = gwnet('cuda', in_dim=2, out_dim=12, kernel_size=3, blocks=1, layers=2, skip_channels=64, end_channels=96,
              supports=[torch.eye(67)] * 2)
    m.skip_convs[0].bias.data.normal_(); m.skip_convs[1].bias.data.normal_()
    pk = m._packed()
    pk.update(pk.pop('head')())
    Wf, Wg = m.filter_convs[1].weight, m.gate_convs[1].weight
    w_fg = pk['w_fg'][1]
    assert w_fg.shape == (3 * 32, 64)
    for (j, c, o) in [(0, 0, 0), (2, 5, 7), (1, 31, 31)]:
        assert w_fg[j * 32 + c, 2 * o] == Wf[o, c, 0, j] and w_fg[j * 32 + c, 2 * o + 1] == Wg[o, c, 0, j]
    assert pk['b_fg'][1][2 * 9] == m.filter_convs[1].bias[9] and pk['b_fg'][1][2 * 9 + 1] == m.gate_convs[1].bias[9]
    Wm = m.gconv[0].mlp.mlp.weight
    assert pk['w_mlp'][0].shape == (224, 32) and pk['w_mlp'][0][100, 3] == Wm[3, 100, 0, 0]
    assert pk['w_skip'].shape == (64, 64) and pk['w_skip'][32 + 4, 17] == m.skip_convs[1].weight[17, 4, 0, 0]
    assert torch.allclose(pk['b_skip'], m.skip_convs[0].bias + m.skip_convs[1].bias)
    assert pk['w_end1'].shape == (64, 96) and pk['w_end1'][5, 70] == m.end_conv_1.weight[70, 5, 0, 0]
    assert pk['w_end2'].shape == (96, 32) and pk['w_end2'][50, 11] == m.end_conv_2.weight[11, 50, 0, 0]
    assert (pk['w_end2'][:, 12:] == 0).all() and (pk['b_end2'][12:] == 0).all()
    # gradient routing: loss = sum_k <packed_k, R_k>  =>  dparam = unpack(R)
    R = {k: ([torch.randn_like(t) for t in v] if isinstance(v, (tuple, list)) else torch.randn_like(v))
         for k, v in pk.items() if k != 'b_mlp'}
    loss = sum((t * r).sum() for k in ('w_fg', 'b_fg') for t, r in zip(pk[k], R[k]))
    loss = loss + (pk['w_mlp'][0] * R['w_mlp'][0]).sum()        # layer 1's mlp unused -> grad must stay None
    for k in ('w_skip', 'b_skip', 'w_end1', 'w_end2', 'b_end2'):
        loss = loss + (pk[k] * R[k]).sum()
    loss.backward()
    assert m.gconv[1].mlp.mlp.weight.grad is None
    assert m.filter_convs[1].weight.grad[7, 5, 0, 2] == R['w_fg'][1][2 * 32 + 5, 14]
    assert m.gate_convs[0].weight.grad[7, 5, 0, 2] == R['w_fg'][0][2 * 32 + 5, 15]
    assert m.gate_convs[1].bias.grad[3] == R['b_fg'][1][7]
    assert m.gconv[0].mlp.mlp.weight.grad[3, 100, 0, 0] == R['w_mlp'][0][100, 3]
    assert m.skip_convs[1].weight.grad[17, 4, 0, 0] == R['w_skip'][36, 17]
    assert torch.equal(m.skip_convs[0].bias.grad, R['b_skip'])
    assert m.end_conv_1.weight.grad[70, 5, 0, 0] == R['w_end1'][5, 70]
    assert m.end_conv_2.weight.grad[11, 50, 0, 0] == R['w_end2'][50, 11]
    assert torch.equal(m.end_conv_2.bias.grad, R['b_end2'][:12])


@pytest.mark.gpu
@pytest.mark.parametrize('V,n_sup,N,Lout,Lin', [(67, 3, 3, 3, 4), (67, 3, 1, 13, 13), (20, 3, 5, 2, 3), (64, 2, 2, 3, 3),
                                                 (68, 3, 2, 2, 4), (33, 1, 3, 3, 3), (5, 3, 7, 1, 2)])
def test_fused_forward_kernel_shapes_against_torch(V, n_sup, N, Lout, Lin):
    """The fused diffusion forward (transposed form where it fits) on shapes that exercise its edges: groups of four
    slabs with a ragged tail, node counts around the 16/32/64 column splits, 1-3 supports.  Reference: the same
    contraction in fp32 torch on the bf16-rounded inputs (u = mlp(concat hops) + b + residual*scale + shift)."""
    from multimodal_outage_b200 import _lib, ops
    lib = _lib.lib()
    torch.manual_seed(V * 100 + n_sup)
    dev, bf = 'cuda', torch.bfloat16
    sups = [torch.softmax(torch.randn(V, V, device=dev), dim=1).contiguous() for _ in range(n_sup)]
    mats = ops.hop_mats(sups)
    H = 2 * n_sup
    z = torch.randn(N, Lout, V, 32, device=dev).to(bf)
    up = torch.randn(N, Lin, V, 32, device=dev).to(bf)
    u = torch.zeros(N, Lout, V, 32, device=dev, dtype=bf)
    w_mlp = (torch.randn(32 * (1 + H), 32, device=dev) / (32 * (1 + H)) ** 0.5).contiguous()
    b_mlp = torch.randn(32, device=dev)
    scale, shift = torch.rand(32, device=dev) + 0.5, torch.randn(32, device=dev)
    ws_w = torch.empty(65536, device=dev, dtype=torch.uint8)
    stats = torch.zeros(64, device=dev, dtype=torch.float64)
    _lib.check(lib.gwn_gcn_fwd(z.data_ptr(), up.data_ptr(), scale.data_ptr(), shift.data_ptr(), mats.data_ptr(), n_sup,
                               w_mlp.data_ptr(), b_mlp.data_ptr(), ws_w.data_ptr(), 0.0, 1, 0, u.data_ptr(), stats.data_ptr(),
                               N, V, Lin, Lout, torch.cuda.current_stream().cuda_stream), 'gwn_gcn_fwd')
    torch.cuda.synchronize()
    zf = z.float()
    cat = [zf]
    for A in sups:                                            # y[w] = sum_v x[v] A[v, w], then once more (graph_wavenet.py:87-93)
        y1 = torch.einsum('nlvc,vw->nlwc', zf, A)
        cat += [y1, torch.einsum('nlvc,vw->nlwc', y1, A)]
    h = torch.cat(cat, dim=-1) @ w_mlp + b_mlp
    ref = h + up.float()[:, Lin - Lout:] * scale + shift
    assert rel(u.float(), ref) < BF16_TOL, rel(u.float(), ref)
    cnt = N * Lout * V
    got_mean = (stats[:32] / cnt).float()
    assert torch.allclose(got_mean, ref.mean(dim=(0, 1, 2)), atol=3e-2), (got_mean - ref.mean(dim=(0, 1, 2))).abs().max()
    got_sq = (stats[32:] / cnt).float()
    assert rel(got_sq, (ref * ref).mean(dim=(0, 1, 2))) < 3e-2


def _gcn_bwd_reference(du, a, b, dz_last, sups, w_mlp, sa, Lf):
    """fp64 restatement of the fused diffusion backward (graph_wavenet.py:76-98 backward + gate backward :222-226) on
    the bf16-rounded inputs; the support gradient comes from autograd of h = sum_j Mt_j-hop(z W_j)."""
    dd = torch.float64
    dh = du.double().cpu()
    af, bfl = a.double().cpu(), b.double().cpu()
    z = (a.float() * b.float()).to(torch.bfloat16).double().cpu()
    W = w_mlp.double().cpu()
    A = [s.double().cpu().clone() for s in sups]
    if sa >= 0:
        A[sa].requires_grad_(True)
    mts = []
    for s in A:
        mts += [s, s @ s]
    dU = [dh] + [torch.einsum('vw,nlwc->nlvc', m.detach(), dh) for m in mts]
    dz = sum(dU[j] @ W[32 * j:32 * j + 32].T for j in range(len(dU)))
    dW = torch.cat([torch.einsum('nlvc,nlvd->cd', z, dU[j]) for j in range(len(dU))], dim=0)
    db = dh.sum(dim=(0, 1, 2))
    if dz_last is not None:
        dz[:, -Lf:] += dz_last.double().cpu()
    df = dz * bfl * (1 - af * af)
    dg = dz * af * bfl * (1 - bfl)
    dfg = torch.stack([df, dg], dim=-1).flatten(-2)
    dA = None
    if sa >= 0:
        h = sum(torch.einsum('vw,nlvc->nlwc', m, z @ W[32 * (j + 1):32 * (j + 2)]) for j, m in enumerate(mts))
        (h * dh).sum().backward()
        dA = A[sa].grad
    return dfg, dW, db, dA


@pytest.mark.gpu
@pytest.mark.parametrize('V,n_sup,N,Lout,Lf,sa', [(67, 3, 3, 3, 1, 2), (67, 3, 2, 13, 13, 2), (67, 3, 5, 1, 1, -1),
                                                  (67, 2, 3, 2, 1, -1), (67, 1, 2, 3, 2, 0), (72, 3, 2, 2, 1, 2),
                                                  (65, 3, 1, 5, 2, 1), (64, 3, 2, 3, 1, 2), (40, 3, 3, 2, 1, 2),
                                                  (33, 2, 3, 3, 1, 1), (20, 3, 5, 2, 2, 2), (5, 3, 7, 1, 1, 2),
                                                  (32, 3, 2, 2, 1, 0)])
def test_transposed_fused_diffusion_backward_kernel(V, n_sup, N, Lout, Lf, sa):
    """The T-form fused diffusion backward (csrc/gcn_fused_bwd_t.cu) through its C-ABI entry on shapes that exercise its
    edges: ragged last group of four slabs, node ranges of 32 / 8 / 24 nodes, 1-3 supports, with and without the
    adaptive-support gradient (delivered factored: dA = d0 + A^T Q + Q A^T), with and without tail rows (dz_last)."""
    from multimodal_outage_b200 import _lib, ops
    lib = _lib.lib()
    if not lib.gwn_gcn_bwd_t_supported(V, n_sup, int(sa >= 0)):
        pytest.skip('no T-form instance for this shape (the node-major kernel runs instead)')
    torch.manual_seed(V * 1000 + n_sup * 10 + N)
    dev, bf = 'cuda', torch.bfloat16
    sups = [torch.softmax(torch.randn(V, V, device=dev), dim=1).contiguous() for _ in range(n_sup)]
    mats = ops.hop_mats(sups)
    H = 2 * n_sup
    du = torch.randn(N, Lout, V, 32, device=dev).to(bf)
    a = torch.tanh(torch.randn(N, Lout, V, 32, device=dev)).to(bf)
    b = torch.sigmoid(torch.randn(N, Lout, V, 32, device=dev)).to(bf)
    dz_last = torch.randn(N, Lf, V, 32, device=dev).to(bf)
    w_mlp = (torch.randn(32 * (1 + H), 32, device=dev) / (32 * (1 + H)) ** 0.5).contiguous()
    dfg = torch.zeros(N, Lout, V, 64, device=dev, dtype=bf)
    dw = torch.empty(32 * (1 + H), 32, device=dev)
    db = torch.empty(32, device=dev)
    dA = torch.zeros(V, V, device=dev)
    dQ = torch.zeros(V, V, device=dev)
    _lib.check(lib.gwn_gcn_bwd_t(du.data_ptr(), a.data_ptr(), b.data_ptr(), dz_last.data_ptr(), mats.data_ptr(), n_sup,
                                 w_mlp.data_ptr(), 0.0, 1, 0, sa, dfg.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                 dA.data_ptr() if sa >= 0 else None, dQ.data_ptr() if sa >= 0 else None, N, V, Lout, Lf,
                                 torch.cuda.current_stream().cuda_stream), 'gwn_gcn_bwd_t')
    torch.cuda.synchronize()
    r_dfg, r_dw, r_db, r_dA = _gcn_bwd_reference(du, a, b, dz_last, sups, w_mlp, sa, Lf)
    assert rel(dfg.float(), r_dfg) < BF16_TOL, rel(dfg.float(), r_dfg)
    assert rel(dw, r_dw) < BF16_TOL, rel(dw, r_dw)
    assert rel(db, r_db) < 1e-2, rel(db, r_db)
    if sa >= 0:
        A = sups[sa].double().cpu()
        full = dA.double().cpu() + A.T @ dQ.double().cpu() + dQ.double().cpu() @ A.T
        assert rel(full, r_dA) < BF16_TOL, rel(full, r_dA)


@pytest.mark.gpu
def test_transposed_fused_backward_draws_the_same_dropout_mask_as_the_node_major_kernel():
    """Same Philox stream in both fused backward kernels: with dropout on, dfg / dW / db agree to accumulation-order
    noise, and the factored support gradient completes to the node-major kernel's dA."""
    from multimodal_outage_b200 import _lib, ops
    lib = _lib.lib()
    V, n_sup, N, Lout, Lf, sa = 67, 3, 6, 3, 1, 2
    torch.manual_seed(5)
    dev, bf = 'cuda', torch.bfloat16
    sups = [torch.softmax(torch.randn(V, V, device=dev), dim=1).contiguous() for _ in range(n_sup)]
    mats = ops.hop_mats(sups)
    du = torch.randn(N, Lout, V, 32, device=dev).to(bf)
    a = torch.tanh(torch.randn(N, Lout, V, 32, device=dev)).to(bf)
    b = torch.sigmoid(torch.randn(N, Lout, V, 32, device=dev)).to(bf)
    dz_last = torch.randn(N, Lf, V, 32, device=dev).to(bf)
    w_mlp = (torch.randn(32 * 7, 32, device=dev) / 15.0).contiguous()
    ws_w = torch.empty(131072, device=dev, dtype=torch.uint8)
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for t_form in (False, True):
        dfg = torch.zeros(N, Lout, V, 64, device=dev, dtype=bf)
        dw, db = torch.empty(32 * 7, 32, device=dev), torch.empty(32, device=dev)
        dA, dQ = torch.zeros(V, V, device=dev), torch.zeros(V, V, device=dev)
        if t_form:
            _lib.check(lib.gwn_gcn_bwd_t(du.data_ptr(), a.data_ptr(), b.data_ptr(), dz_last.data_ptr(), mats.data_ptr(),
                                         n_sup, w_mlp.data_ptr(), 0.3, 1234, 7, sa, dfg.data_ptr(), dw.data_ptr(),
                                         db.data_ptr(), dA.data_ptr(), dQ.data_ptr(), N, V, Lout, Lf, st), 'gwn_gcn_bwd_t')
        else:
            _lib.check(lib.gwn_gcn_bwd(du.data_ptr(), a.data_ptr(), b.data_ptr(), dz_last.data_ptr(), mats.data_ptr(),
                                       n_sup, w_mlp.data_ptr(), ws_w.data_ptr(), 0.3, 1234, 7, sa, dfg.data_ptr(),
                                       dw.data_ptr(), db.data_ptr(), dA.data_ptr(), N, V, Lout, Lf, st), 'gwn_gcn_bwd')
        torch.cuda.synchronize()
        A = sups[sa]
        out.append((dfg.float(), dw, db, dA + A.T @ dQ + dQ @ A.T))
    for x, y, tol in zip(out[1], out[0], (1e-2, 1e-2, 5e-3, 1e-2)):
        assert rel(x, y) < tol, rel(x, y)


# ------------------------------------------------------------------------------------------------------------------
# bf16 tensor-core path at the shapes the benchmarked / reference configurations use beyond dilation 1-2 and k = 2
@pytest.mark.parametrize('dil,Lin', [(4, 7), (8, 11)])
def test_bf16_layer_op_config5_dilations(dil, Lin):
    """Config 5's layers run dilations 4 and 8 (4x4 layers): every output and gradient of the layer op within 2e-2 of
    the fp64 oracle on the tensor-core path - supports resident on chip (V=67, with and without the adaptive-support
    gradient = both fused backward kernels) and the TMA-tiled big-graph path (V=150)."""
    for kw in (dict(V=67, N=3, adp_grad=True, seed=31), dict(V=67, N=2, adp_grad=False, seed=32),
               dict(V=150, N=2, adp_grad=True, seed=33)):
        errs = _layer_op_case(torch.bfloat16, BF16_TOL, Lin=Lin, dil=dil, taps=2, n_sup=3, with_bn=True, mask=True,
                              tensor_cores=True, **kw)
        print(f'bf16 layer op, dilation {dil}:', kw, {k: f'{v:.1e}' for k, v in errs.items()})


@pytest.mark.parametrize('V,n_sup', [(67, 2), (67, 3), (150, 2)])
def test_bf16_layer_op_kernel_size_1(V, n_sup):
    """kernel_size = 1 is the reference's literal default (graph_wavenet.py:101): the gated conv is a plain 1x1 pair (one
    chunk, Lout = Lin), nothing is cropped from the residual.  Tensor-core path, 2e-2 on everything."""
    for with_bn, mask in ((True, False), (False, True)):
        errs = _layer_op_case(torch.bfloat16, BF16_TOL, V=V, N=3, Lin=3, dil=1, taps=1, n_sup=n_sup, with_bn=with_bn,
                              mask=mask, seed=41, tensor_cores=True)
        print(f'bf16 layer op, kernel size 1, V={V}:', {k: f'{v:.1e}' for k, v in errs.items()})


def test_literal_reference_call_in_bf16():
    """The literal `[67, h, 320]` call of Modified_UNET.forward (unet.py:224-226; kernel_size 1, supports [I] + adaptive,
    batch 1) under bf16: output and loss within 2e-2 of what the REFERENCE's own classes produced in fp32
    (tests/golden/literal.npz); gradients against the exact oracle, fixed bars as in `_oracle_case`."""
    from multimodal_outage_b200 import ops
    c = GOLDEN_CASES['literal']
    cfg, g = c['cfg'], _golden('literal')
    sup = case_supports(c['supports'])
    m = build_model(cfg, sup, horizon=c['horizon'])
    sd = load_synth(m, cfg, c['seed'])
    x_np, _ = case_inputs('literal')
    x = torch.tensor(x_np, device='cuda', requires_grad=True)
    assert x.dim() == 3
    m.train()
    m.compute_dtype = torch.bfloat16
    ops.HEAD_CAPTURE = {}
    try:
        out = m(x)
    finally:
        cap, ops.HEAD_CAPTURE = ops.HEAD_CAPTURE, None
    assert tuple(out.shape) == g['out_train'].shape and out.dtype == torch.float32
    assert rel(out, g['out_train']) < BF16_TOL, rel(out, g['out_train'])
    loss = torch.nn.functional.mse_loss(out, torch.tensor(g['target'], device='cuda'))
    assert abs(loss.item() - float(g['loss'])) < BF16_TOL * abs(float(g['loss']))
    loss.backward()
    h = c['horizon']
    m1, m2 = captured_head_masks(cap, 1, 67, h)
    _, _, grads_p, _ = oracle_run(cfg, sd, x_np, sup, g['target'], literal=True, horizon=h, head_masks=(m1, m2))
    worst = {}
    for k, p in [('__x__', x)] + list(m.named_parameters()):
        gr = grads_p.get(k)
        if gr is None or p.grad is None or (k.endswith('bias') and 'gconv' in k):
            continue
        worst[k] = rel(p.grad, gr.reshape(p.grad.shape))
    bad = {k: v for k, v in worst.items() if not v < BF16_TOL}
    print('literal bf16: worst gradient (head decisions pinned)', max(worst.values()), 'x.grad vs reference fp32',
          rel(x.grad, g['x_grad']))
    assert not bad, bad
    assert rel(x.grad, g['x_grad']) < 0.12
    m.eval()
    with torch.no_grad():
        assert rel(m(x.detach()), g['out_eval']) < BF16_TOL


def test_hop_and_support_gradient_at_3100_nodes():
    """The 3,100-node configurations (BASELINE configs 3 and 5) run `gwn_hop_big` / `gwn_dadj_big` at V = 3100: one hop
    through each operand image (A and A^T) and the support gradient against fp64 einsums, at the real graph size
    (ragged: 3100 = 24 x 128 + 28) and a slab count that is not a multiple of the 8-slab box."""
    from multimodal_outage_b200 import ops, _lib
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(2)
    V, slabs = 3100, 11
    A = torch.softmax(torch.randn(V, V, device='cuda') * 3, dim=1)
    img = ops.support_images([A])
    x = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
    g = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
    y = torch.empty_like(x)
    Ab = A.to(torch.bfloat16).double()
    for which in range(2):
        ref = torch.einsum('vw,svc->swc' if which == 0 else 'wv,svc->swc', Ab, x.double())
        _lib.check(lib.gwn_hop_big(img.data_ptr(), 1, 0, which, x.data_ptr(), y.data_ptr(), None, slabs, V, st), 'hop')
        assert rel(y, ref) < 4e-3, (which, rel(y, ref))                 # bf16 output rounding only
        _lib.check(lib.gwn_hop_big(img.data_ptr(), 1, 0, which, x.data_ptr(), y.data_ptr(), g.data_ptr(), slabs, V, st), 'hop')
        assert rel(y, ref + g.double()) < 4e-3
    dA = torch.zeros(V, V, device='cuda')
    _lib.check(lib.gwn_dadj_big(x.data_ptr(), g.data_ptr(), dA.data_ptr(), slabs, V, st), 'dadj')
    ref = torch.einsum('svc,swc->vw', x.double(), g.double())
    assert rel(dA, ref) < 1e-5, rel(dA, ref)


@pytest.mark.parametrize('V,slabs', [(400, 80), (3100, 100)])
def test_support_gradient_split_k_on_cta_pairs(V, slabs):
    """`gwn_dadj_big` at slab counts where the launcher cuts the contraction into several parts (split-K over the slab
    range, partial sums added with vector reductions) on the two-SM kernel: accumulates INTO dA, within 1e-5 of the fp64
    einsum; and the two-SM hop at a slab count that is not a multiple of its 8-slab tile."""
    from multimodal_outage_b200 import ops, _lib
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(4)
    x = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
    g = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
    dA = torch.full((V, V), 0.5, device='cuda')
    for _ in range(2):                                                       # two accumulating calls
        _lib.check(lib.gwn_dadj_big(x.data_ptr(), g.data_ptr(), dA.data_ptr(), slabs, V, st), 'dadj')
    ref = 0.5 + 2.0 * torch.einsum('svc,swc->vw', x.double(), g.double())
    assert rel(dA, ref) < 1e-5, (V, slabs, rel(dA, ref))
    A = torch.softmax(torch.randn(V, V, device='cuda') * 2, dim=1)
    img = ops.support_images([A])
    y = torch.empty_like(x)
    refh = torch.einsum('vw,svc->swc', A.to(torch.bfloat16).double(), x.double())
    _lib.check(lib.gwn_hop_big(img.data_ptr(), 1, 0, 0, x.data_ptr(), y.data_ptr(), None, slabs, V, st), 'hop')
    assert rel(y, refh) < 4e-3, (V, slabs, rel(y, refh))


def test_fused_dropout_statistics():
    """The in-kernel Philox stream (7 rounds, one byte per element): realised drop rate, per-channel and per-node
    uniformity and lag-1 independence over 1.1e7 draws, at the benchmark's p = 0.3 (realised as 77/256)."""
    from multimodal_outage_b200 import ops, _lib
    lib = _lib.lib()
    N, L, V = 512, 10, 67
    n = N * L * V * 32
    ones = torch.ones(N, L, V, 32, device='cuda', dtype=torch.bfloat16)
    out = torch.empty_like(ones)
    _lib.check(lib.gwn_dropout_apply(ones.data_ptr(), out.data_ptr(), n // 32, 0.3, 12345, 7,
                                     torch.cuda.current_stream().cuda_stream), 'gwn_dropout_apply')
    keep = (out != 0)
    p_real = 77 / 256
    rate = 1.0 - keep.double().mean().item()
    sigma = (p_real * (1 - p_real) / n) ** 0.5
    assert abs(rate - p_real) < 5 * sigma, (rate, p_real, sigma)
    kept_val = out[keep].float()
    assert torch.allclose(kept_val, torch.full_like(kept_val, 256 / (256 - 77)), rtol=4e-3)    # unbiased for the realised rate
    per_c = 1.0 - keep.double().mean(dim=(0, 1, 2))
    per_v = 1.0 - keep.double().mean(dim=(0, 1, 3))
    assert (per_c - p_real).abs().max().item() < 5 * (p_real * (1 - p_real) / (n / 32)) ** 0.5
    assert (per_v - p_real).abs().max().item() < 5 * (p_real * (1 - p_real) / (n / V)) ** 0.5
    k = keep.double().flatten()
    for lag in (1, 16, 32, 32 * V):                     # neighbouring channel, next Philox call, next node, next time step
        a, b = k[:-lag] - (1 - p_real), k[lag:] - (1 - p_real)
        corr = (a * b).mean().item() / (p_real * (1 - p_real))
        assert abs(corr) < 5 / (n - lag) ** 0.5, (lag, corr)


def test_bf16_config2_full_batch_training_parity():
    """BASELINE config 2 exactly as benchmarked - 67-county graph, fwd/bwd/adaptive supports, k = 2, 4x2 layers, batch 512,
    training mode with dropout masks p = 0.3 (explicit, so the fp64 oracle applies the same masks) - through the same
    fixed bars as every whole-model bf16 case: output / loss 2e-2, every gradient 2e-2 with the head's ReLU decisions
    pinned, 0.12 unpinned."""
    cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, dropout=0.3)
    sup = double_transition(np.load(os.path.join(GOLDEN_DIR, 'adj_mx_fl.npy')).astype(np.float32))
    _oracle_case(cfg, sup, n=512, t_in=12, seed=12, dtype=torch.bfloat16, tol=BF16_TOL, masks=True)


def test_bf16_layer_op_at_3100_nodes():
    """One layer of the 3,100-node configurations (BASELINE configs 3 and 5) at the REAL graph size: TMA-tiled hop GEMMs,
    the Horner-form backward and the support gradient `gwn_dadj_big`, every output and gradient within 2e-2 of the fp64
    oracle (one sample, three time steps - the oracle's 3100 x 3100 products stay in the seconds)."""
    errs = _layer_op_case(torch.bfloat16, BF16_TOL, V=3100, N=1, Lin=3, dil=1, taps=2, n_sup=3, with_bn=True, mask=True,
                          seed=51, tensor_cores=True)
    print('bf16 layer op at V=3100:', {k: f'{v:.1e}' for k, v in errs.items()})


def test_sparse_support_hop_matches_the_dense_hop():
    """`gwn_hop_ell` (ELL gather) against an fp64 einsum and against the dense tensor-core hop, both directions, with and
    without the add-in (in place), on the transition matrices of a 3,100-node kNN graph (<= 20 non-zeros per row)."""
    from multimodal_outage_b200 import ops, _lib
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(3)
    V, slabs = 3100, 7
    A = torch.tensor(np.asarray(double_transition(synthetic_knn_graph(V))[0]), dtype=torch.float32, device='cuda').contiguous()
    assert ops.register_sparse_support(A)
    e = ops._ELL_REGISTRY[A.data_ptr()]
    assert e['width'] <= ops.ELL_MAX_WIDTH and int((A != 0).sum(dim=1).max()) <= e['width']
    x = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
    add = torch.randn(slabs, V, 32, device='cuda').to(torch.bfloat16)
    img = ops.support_images([A])
    for which in range(2):
        ref = torch.einsum('vw,svc->swc' if which == 0 else 'wv,svc->swc', A.double(), x.double())
        y = torch.empty_like(x)
        _lib.check(lib.gwn_hop_ell(e['idx'][which].data_ptr(), e['val'][which].data_ptr(), e['width'], x.data_ptr(), y.data_ptr(),
                                   None, slabs, V, st), 'hop_ell')
        assert rel(y, ref) < 4e-3, (which, rel(y, ref))                      # bf16 output rounding only
        yd = torch.empty_like(x)
        _lib.check(lib.gwn_hop_big(img.data_ptr(), 1, 0, which, x.data_ptr(), yd.data_ptr(), None, slabs, V, st), 'hop')
        assert rel(y, yd) < 8e-3
        y2 = add.clone()                                                      # y += A x, in place
        _lib.check(lib.gwn_hop_ell(e['idx'][which].data_ptr(), e['val'][which].data_ptr(), e['width'], x.data_ptr(), y2.data_ptr(),
                                   y2.data_ptr(), slabs, V, st), 'hop_ell')
        assert rel(y2, ref + add.double()) < 4e-3
    # a modified support is no longer served from the registry
    A.mul_(1.0)
    assert ops._ell_array([A])[0] is None


@pytest.mark.parametrize('V,N', [(150, 3), (333, 2)])
def test_bf16_layer_op_big_graph_sparse_fixed_supports(V, N):
    """The big-graph layer op with its two fixed supports applied as sparse gathers (forward hops, Horner backward,
    recompute backward share `hop_any`) and the adaptive one as a dense GEMM: everything within 2e-2 of the fp64 oracle."""
    for with_bn, mask in ((True, False), (False, True)):
        errs = _layer_op_case(torch.bfloat16, BF16_TOL, V=V, N=N, Lin=5, dil=1, taps=2, n_sup=3, with_bn=with_bn, mask=mask,
                              seed=61, tensor_cores=True, sparse=True)
        print(f'bf16 big-graph layer op, sparse fixed supports, V={V}:', {k: f'{v:.1e}' for k, v in errs.items()})


@pytest.mark.parametrize('shape', [(512, 12, 67, 1), (3, 5, 7), (4099,), (1,)])
def test_fused_mse_loss_matches_torch(shape):
    """`ops.mse_loss` (nn.MSELoss of the training step, lit.py:24: one cluster launch forward, one launch backward) against
    torch.nn.functional.mse_loss: value and gradient to fp32 rounding, incl. sizes that are not multiples of four, an
    unaligned view, a non-unit upstream gradient, and CUDA-graph replay (no workspace to re-zero)."""
    from multimodal_outage_b200 import ops
    torch.manual_seed(5)
    a = torch.randn(*shape, device='cuda', requires_grad=True)
    b = torch.randn(*shape, device='cuda')
    l0 = torch.nn.functional.mse_loss(a, b)
    (g0,) = torch.autograd.grad(l0 * 3.0, a)
    l1 = ops.mse_loss(a, b)
    (g1,) = torch.autograd.grad(l1 * 3.0, a)
    assert abs(l1.item() - l0.item()) <= 2e-6 * abs(l0.item()) + 1e-12, (l1.item(), l0.item())
    assert rel(g1, g0) < 1e-6
    if len(shape) == 1 and shape[0] > 8:                       # 4-byte aligned views: the scalar path
        av, bv = a.detach()[1:], b[1:]
        assert abs(ops.mse_loss(av.contiguous(), bv.contiguous()).item() - torch.nn.functional.mse_loss(av, bv).item()) < 1e-5
        from multimodal_outage_b200 import _lib
        out = torch.empty((), device='cuda')
        _lib.check(_lib.lib().gwn_mse_loss_fwd(av.data_ptr(), bv.data_ptr(), av.numel(), out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), 'mse')
        assert abs(out.item() - torch.nn.functional.mse_loss(av, bv).item()) < 1e-5
    if shape == (512, 12, 67, 1):
        ad = a.detach()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            ops.mse_loss(ad, b)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            lg = ops.mse_loss(ad, b)
        for _ in range(3):
            ad.add_(0.25)
            graph.replay()
            assert abs(lg.item() - torch.nn.functional.mse_loss(ad, b).item()) < 1e-5
