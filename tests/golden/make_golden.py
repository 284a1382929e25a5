#!/usr/bin/env python
"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE.

Run in the build container only (needs /root/reference; the GPU box does not have
it):  ``python tests/golden/make_golden.py``.

How the reference is run: ``models/graph_wavenet.py`` cannot be imported (it pulls
matplotlib, reads a hard-coded /home path and moves tensors to CUDA at import
time), so the four ClassDefs ``nconv/linear/gcn/gwnet`` (lines
60/68/76/100) are AST-extracted and exec'd verbatim with the module globals they
need.  "general" mode drops exactly the two statements at source lines 189 and 255
(the hard-coded ``.view``s); "literal" mode keeps everything.  ``asym_adj`` is
extracted from ``utils.py:152``.  No reference source is copied into the repo.

Weights are NOT stored: both this script and the tests derive them from
``oracle.gwnet_oracle.synthetic_state_dict(cfg, seed)`` (numpy PCG64).
"""
import ast
import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))
from oracle.gwnet_oracle import GWNetConfig, synthetic_state_dict          # noqa: E402
from oracle.graph_oracle import synthetic_directed_graph                   # noqa: E402
from oracle.cases import GOLDEN_CASES, case_supports, case_inputs          # noqa: E402

REF = '/root/reference'
GW_SRC = os.path.join(REF, 'models', 'graph_wavenet.py')
UT_SRC = os.path.join(REF, 'utils.py')
EXPECTED_CLASS_LINES = {'nconv': 60, 'linear': 68, 'gcn': 76, 'gwnet': 100}


def load_reference_classes(general: bool, default_supports):
    src = open(GW_SRC).read()
    tree = ast.parse(src)
    body = []
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in EXPECTED_CLASS_LINES:
            assert node.lineno == EXPECTED_CLASS_LINES[node.name], (node.name, node.lineno)
            if node.name == 'gwnet' and general:
                for fn in node.body:
                    if isinstance(fn, ast.FunctionDef) and fn.name == 'forward':
                        fn.body = [s for s in fn.body if s.lineno not in (189, 255)]
            body.append(node)
    glb = {'torch': torch, 'nn': nn, 'F': F, 'n_counties': 67, 'feature_vector_size': 256,
           'time_embed_size': 64, 'default_supports': default_supports}
    exec(compile(ast.Module(body=body, type_ignores=[]), GW_SRC, 'exec'), glb)
    return glb


def load_reference_asym_adj():
    import scipy.sparse as sp
    tree = ast.parse(open(UT_SRC).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'asym_adj']
    assert fn and fn[0].lineno == 152
    glb = {'np': np, 'sp': sp}
    exec(compile(ast.Module(body=fn, type_ignores=[]), UT_SRC, 'exec'), glb)
    return glb['asym_adj']


def fl_adjacency():
    import pandas as pd
    df = pd.read_csv(os.path.join(REF, 'data', 'graph', 'adj_mx_fl.csv'), index_col=0)
    return df.values.astype(np.int64)


def grad_summary(t: torch.Tensor):
    f = t.detach().double().flatten()
    return np.array([f.norm().item(), f.sum().item(), f.abs().sum().item()], dtype=np.float64), \
        f[:16].numpy().astype(np.float64)


def run_case(name, cfg: GWNetConfig, *, supports_np, n, t_in, seed, literal=False, horizon=None,
             full_grads=False, gcn_bool=True):
    sup = [torch.tensor(s) for s in supports_np]
    glb = load_reference_classes(general=not literal, default_supports=sup)
    torch.manual_seed(0)
    model = glb['gwnet']('cpu', num_nodes=cfg.num_nodes, dropout=cfg.dropout,
                         supports=sup if cfg.n_fixed_supports else None,
                         gcn_bool=gcn_bool, addaptadj=cfg.adaptive, in_dim=cfg.in_dim,
                         out_dim=cfg.out_dim, horizon=horizon or 1,
                         residual_channels=cfg.residual_channels,
                         dilation_channels=cfg.dilation_channels, skip_channels=cfg.skip_channels,
                         end_channels=cfg.end_channels, kernel_size=cfg.kernel_size,
                         blocks=cfg.blocks, layers=cfg.layers)
    sd = synthetic_state_dict(cfg, seed)
    ref_keys = list(model.state_dict().keys())
    missing = model.load_state_dict({k: v for k, v in sd.items() if k in ref_keys}, strict=True)
    x_np, rng = case_inputs(name)
    x = torch.tensor(x_np)
    x.requires_grad_(True)
    model.train()
    out = model(x)
    y = torch.tensor(rng.standard_normal(tuple(out.shape)), dtype=torch.float32)
    loss = F.mse_loss(out, y)
    loss.backward()
    rec = {'out_train': out.detach().numpy(), 'target': y.numpy(), 'loss': np.float64(loss.item()),
           'x_grad': x.grad.numpy(), 'state_keys': np.array(ref_keys)}
    for k, p in model.named_parameters():
        if p.grad is None:
            rec[f'gradnone/{k}'] = np.zeros(0)
            continue
        s, head = grad_summary(p.grad)
        rec[f'gradsum/{k}'] = s
        rec[f'gradhead/{k}'] = head
        if full_grads:
            rec[f'grad/{k}'] = p.grad.numpy()
    for k, b in model.named_buffers():
        rec[f'buf/{k}'] = b.detach().numpy()
    model.eval()
    with torch.no_grad():
        rec['out_eval'] = model(x.detach()).numpy()
    path = os.path.join(HERE, f'{name}.npz')
    np.savez_compressed(path, **rec)
    print(f'{name}: out {tuple(out.shape)} loss {loss.item():.6f} -> {os.path.getsize(path)/1e3:.0f} kB')


def main():
    # fingerprint of the reference sources the goldens were generated from
    fp = {os.path.basename(p): hashlib.sha256(open(p, 'rb').read()).hexdigest() for p in (GW_SRC, UT_SRC)}
    adj_fl = fl_adjacency()
    assert adj_fl.shape == (67, 67) and (adj_fl == adj_fl.T).all() and adj_fl.sum() == 312
    np.save(os.path.join(HERE, 'adj_mx_fl.npy'), adj_fl.astype(np.int8))

    asym = load_reference_asym_adj()
    dirg = synthetic_directed_graph(67)
    dirg_big = synthetic_directed_graph(301, p=0.02, seed=11)
    np.savez_compressed(
        os.path.join(HERE, 'asym_adj.npz'),
        fl_f32=np.asarray(asym(adj_fl.astype(np.float32))),
        fl_f64=np.asarray(asym(adj_fl.astype(np.float64))),
        dir_fwd=np.asarray(asym(dirg.astype(np.float32))),
        dir_bwd=np.asarray(asym(dirg.T.astype(np.float32))),
        big_fwd=np.asarray(asym(dirg_big.astype(np.float32))),
        sha=np.array([f'{k}:{v}' for k, v in fp.items()]))
    try:
        asym(adj_fl)
        raise SystemExit('reference asym_adj unexpectedly accepted int input')
    except ValueError:
        pass

    for name, c in GOLDEN_CASES.items():
        run_case(name, c['cfg'], supports_np=case_supports(c['supports'], asym=asym), n=c['n'],
                 t_in=c['t_in'], seed=c['seed'], literal=c.get('literal', False),
                 horizon=c.get('horizon'), full_grads=c.get('full_grads', False),
                 gcn_bool=c['cfg'].gcn_bool)


if __name__ == '__main__':
    main()
