"""Whole-model bf16 gradient error table: per parameter class, relative L2 error of the CUDA bf16
path against the exact fp64 oracle, next to (i) the bf16-storage oracle's own error, (ii) the same comparison with the
head's ReLU DECISIONS pinned to the CUDA run's (a ReLU net is piecewise linear: with the decisions equal, the remaining
error is the smooth part), (iii) the storage oracle with the residual stream kept unrounded (what an fp32 `u` / `dx`
would buy).  Writes profiles/r2_bf16_grad_parity.json and prints a markdown table."""
import json, os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import oracle.gwnet_oracle as go
from oracle.cases import GOLDEN_DIR, case_supports
from oracle.graph_oracle import double_transition, synthetic_knn_graph
from gpu_helpers import build_model, load_synth, oracle_run, rel, captured_head_masks
from multimodal_outage_b200 import ops

torch.set_num_threads(os.cpu_count() or 8)
GROUPS = ['x', 'start_conv', 'filter_convs', 'gate_convs', 'gconv.weight', 'bn.weight', 'bn.bias', 'skip_convs', 'nodevec',
          'end_conv_1', 'end_conv_2']


def group_of(k):
    if k == '__x__': return 'x'
    if k.startswith('gconv'): return 'gconv.weight' if k.endswith('weight') else None      # bias grads are analytically 0
    if k.startswith('bn.'): return 'bn.weight' if k.endswith('weight') else 'bn.bias'
    for g in ('start_conv', 'filter_convs', 'gate_convs', 'skip_convs', 'nodevec', 'end_conv_1', 'end_conv_2'):
        if k.startswith(g): return g
    return None


def worst(ga, gb):
    """worst per-tensor relative L2 per group; conv biases feeding a training BN are analytically zero -> skipped"""
    out = {}
    for k, b in gb.items():
        a = ga.get(k)
        g = group_of(k)
        if g is None or a is None or b is None: continue
        if k.endswith('bias') and (k.startswith('filter_convs') or k.startswith('gate_convs') or k.startswith('start_conv')):
            pass
        out[g] = max(out.get(g, 0.0), rel(a, b))
    return out


def run_case(name, cfg, sup, n, t_in, seed, masks):
    t0 = time.time()
    m = build_model(cfg, sup)
    sd = load_synth(m, cfg, seed)
    rng = np.random.default_rng(seed + 1)
    x_np = rng.standard_normal((n, cfg.in_dim, cfg.num_nodes, t_in)).astype(np.float32)
    L = go.layer_lengths(cfg, t_in)
    y_np = rng.standard_normal((n, cfg.out_dim, cfg.num_nodes, L[-1])).astype(np.float32)
    dm_o = dm_g = None
    if masks:
        keep = 1.0 - cfg.dropout
        dm_np = [(rng.random((n, 32, cfg.num_nodes, L[i + 1])) < keep).astype(np.float32) / keep for i in range(cfg.n_layers)]
        dm_o = [torch.tensor(d, dtype=torch.float64) for d in dm_np]
        dm_g = [torch.tensor(d, device='cuda') for d in dm_np]
    # CUDA bf16
    x = torch.tensor(x_np, device='cuda', requires_grad=True)
    m.train(); m.compute_dtype = torch.bfloat16
    ops.HEAD_CAPTURE = {}
    out = m(x, dropout_masks=dm_g)
    cap, ops.HEAD_CAPTURE = ops.HEAD_CAPTURE, None
    loss = torch.nn.functional.mse_loss(out, torch.tensor(y_np, device='cuda'))
    loss.backward()
    g_cuda = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    g_cuda['__x__'] = x.grad
    out_o, loss_o, g_exact, tr = oracle_run(cfg, sd, x_np, sup, y_np, dropout_masks=dm_o)
    m1, m2 = captured_head_masks(cap, n, cfg.num_nodes, L[-1])
    flips1 = float(((tr.skip.detach() > 0).double() != m1).double().mean())
    _, _, g_pinned, _ = oracle_run(cfg, sd, x_np, sup, y_np, dropout_masks=dm_o, head_masks=(m1, m2))
    _, _, g_store, _ = oracle_run(cfg, sd, x_np, sup, y_np, dropout_masks=dm_o, storage=torch.bfloat16)
    go.STORAGE_EXCLUDE = {'u'}
    _, _, g_store_u32, _ = oracle_run(cfg, sd, x_np, sup, y_np, dropout_masks=dm_o, storage=torch.bfloat16)
    go.STORAGE_EXCLUDE = set()
    rec = {'case': name, 'N': n, 'V': cfg.num_nodes, 'T': t_in, 'layers': f'{cfg.blocks}x{cfg.layers}', 'dropout_masks': bool(masks),
           'out_rel': rel(out, out_o), 'loss_rel': abs(loss.item() - loss_o.item()) / abs(loss_o.item()),
           'head_relu1_decisions_flipped': flips1,
           'cuda_vs_exact': worst(g_cuda, g_exact), 'cuda_vs_exact_head_decisions_pinned': worst(g_cuda, g_pinned),
           'storage_oracle_vs_exact': worst(g_store, g_exact), 'storage_oracle_fp32_residual_vs_exact': worst(g_store_u32, g_exact),
           'seconds': time.time() - t0}
    print(f'## {name}: N={n} V={cfg.num_nodes} T={t_in} {cfg.blocks}x{cfg.layers} layers; out {rec["out_rel"]:.2e}, loss {rec["loss_rel"]:.1e}, '
          f'{100 * flips1:.2f}% of the first head ReLU decisions differ from the exact oracle ({rec["seconds"]:.0f} s)')
    print('| gradient of | CUDA bf16 vs exact | ... head ReLU decisions pinned | bf16-storage oracle vs exact | ... with fp32 residual stream |')
    print('|---|---|---|---|---|')
    for g in GROUPS:
        if g in rec['cuda_vs_exact']:
            print(f'| {g} | {rec["cuda_vs_exact"][g]:.2e} | {rec["cuda_vs_exact_head_decisions_pinned"].get(g, float("nan")):.2e} | '
                  f'{rec["storage_oracle_vs_exact"].get(g, float("nan")):.2e} | {rec["storage_oracle_fp32_residual_vs_exact"].get(g, float("nan")):.2e} |')
    sys.stdout.flush()
    return rec


if __name__ == '__main__':
    big = int(os.environ.get('TABLE_N', '512'))
    fl = double_transition(np.load(os.path.join(GOLDEN_DIR, 'adj_mx_fl.npy')).astype(np.float32))
    C = go.GWNetConfig
    recs = [
        run_case('c1-shape (directed supports)', C(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, dropout=0.0), case_supports('dir'), 8, 12, 11, False),
        run_case('config 5 structure (4x4 layers, dilation <= 8)', C(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=4, layers=4, dropout=0.0), case_supports('dir'), 4, 48, 21, False),
        run_case('V=310 kNN graph (TMA-tiled hop path)', C(num_nodes=310, in_dim=2, out_dim=12, kernel_size=2, blocks=2, layers=2, skip_channels=64, end_channels=128, dropout=0.0),
                 double_transition(synthetic_knn_graph(310)), 2, 12, 15, False),
        run_case(f'config 2, training mode, explicit dropout masks p=0.3', C(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, dropout=0.3), fl, big, 12, 12, True),
    ]
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(recs, open(os.path.join(ROOT, 'gpurun_out', 'r2_bf16_grad_parity.json'), 'w'), indent=1)
