"""Shared helpers for the -m gpu parity tests (oracle = checker only)."""
import numpy as np
import torch

from oracle.cases import GOLDEN_CASES, case_inputs, case_supports
from oracle.gwnet_oracle import ForwardTrace, gwnet_forward, gwnet_forward_literal, synthetic_state_dict


def build_model(cfg, supports_np, device='cuda', horizon=1):
    from multimodal_outage_b200 import gwnet
    sup = [torch.tensor(s) for s in supports_np]
    m = gwnet(device, num_nodes=cfg.num_nodes, dropout=cfg.dropout,
              supports=sup if cfg.n_fixed_supports else None, gcn_bool=cfg.gcn_bool, addaptadj=cfg.adaptive,
              in_dim=cfg.in_dim, out_dim=cfg.out_dim, horizon=horizon, residual_channels=cfg.residual_channels,
              dilation_channels=cfg.dilation_channels, skip_channels=cfg.skip_channels,
              end_channels=cfg.end_channels, kernel_size=cfg.kernel_size, blocks=cfg.blocks, layers=cfg.layers)
    return m


def load_synth(model, cfg, seed):
    sd = synthetic_state_dict(cfg, seed)
    model.load_state_dict({k: v for k, v in sd.items() if k in model.state_dict()}, strict=True)
    return sd


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu().flatten()
    b = torch.as_tensor(b).detach().double().cpu().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def oracle_run(cfg, sd, x_np, sup_np, target_np=None, literal=False, horizon=None, dtype=torch.float64,
               dropout_masks=None, training=True, storage=None, head_masks=None):
    """Runs the CPU oracle (fp64 by default); returns out, loss, grads dict, trace."""
    sdo = {}
    for k, v in sd.items():
        t = v.clone().to(dtype) if v.is_floating_point() else v.clone()
        if t.is_floating_point() and 'running' not in k:
            t.requires_grad_(True)
        sdo[k] = t
    sup = [torch.tensor(s).to(dtype) for s in sup_np]
    x = torch.tensor(x_np).to(dtype).requires_grad_(True)
    tr = ForwardTrace()
    if literal:
        out = gwnet_forward_literal(sdo, x, sup, cfg, horizon, training=training, trace=tr,
                                    dropout_masks=dropout_masks, storage=storage, head_masks=head_masks)
    else:
        out = gwnet_forward(sdo, x, sup, cfg, training=training, trace=tr, dropout_masks=dropout_masks,
                            storage=storage, head_masks=head_masks)
    loss = None
    grads = {}
    if target_np is not None and training:
        loss = torch.nn.functional.mse_loss(out, torch.tensor(target_np).to(dtype))
        loss.backward()
        grads = {k: v.grad for k, v in sdo.items() if v.is_floating_point() and v.requires_grad}
        grads['__x__'] = x.grad
    return out.detach(), loss, grads, tr


def compare_grads(model, grads, tol, report=None):
    """Per-tensor relative L2; analytically-zero bias grads use atol scaled to the
    same layer's weight-grad norm.  Returns list of (name, err) failures."""
    bad = []
    wn = {k: (float(g.norm()) if g is not None else 0.0) for k, g in grads.items()}
    for k, p in model.named_parameters():
        g_ref = grads.get(k)
        if g_ref is None:
            if p.grad is not None and float(p.grad.abs().max()) != 0.0:
                bad.append((k, 'expected None/zero grad'))
            continue
        if p.grad is None:
            bad.append((k, 'grad is None'))
            continue
        diff = (p.grad.detach().double().cpu() - g_ref.double()).norm().item()
        scale = wn[k]
        if k.endswith('bias'):
            scale = max(scale, wn.get(k[:-4] + 'weight', 0.0))
        err = diff / max(scale, 1e-30)
        if report is not None:
            report.append((k, err))
        if not err <= tol:
            bad.append((k, err))
    return bad


def captured_head_masks(cap, n, v, lf):
    """ReLU decisions of the CUDA head (ops.HEAD_CAPTURE after a bf16 forward) as oracle-layout 0/1 masks
    ([N,S,V,Lf], [N,E,V,Lf]).  s1 is saved as [P, hi | lo] of relu(skip), e1 as relu(end_conv_1)."""
    s1, e1 = cap['s1'].float(), cap['e1'].float()
    S = s1.shape[1] // 2
    m1 = ((s1[:, :S] != 0) | (s1[:, S:] != 0)).view(n, lf, v, S).permute(0, 3, 2, 1).double().cpu()
    m2 = (e1 != 0).view(n, lf, v, -1).permute(0, 3, 2, 1).double().cpu()
    return m1, m2
