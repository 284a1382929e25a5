"""Layer-by-layer comparison of the bf16 CUDA path with the storage-rounding oracle."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', 'tests')))
import numpy as np, torch
from oracle.cases import case_supports
from oracle.gwnet_oracle import GWNetConfig, ForwardTrace, gwnet_forward, synthetic_state_dict
from gpu_helpers import build_model, load_synth, rel
from multimodal_outage_b200 import ops

cfg = GWNetConfig(num_nodes=67, in_dim=2, out_dim=12, kernel_size=2, blocks=2, layers=2, skip_channels=64, end_channels=128, dropout=0.0)
sup = case_supports('dir')
m = build_model(cfg, sup); sd = load_synth(m, cfg, 7)
rng = np.random.default_rng(8)
x_np = rng.standard_normal((4, 2, 67, 12)).astype(np.float32)
sdo = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
for storage in (None, torch.bfloat16):
    tr = ForwardTrace()
    ref = gwnet_forward(sdo, torch.tensor(x_np).double(), [torch.tensor(s).double() for s in sup], cfg, training=True, trace=tr, storage=storage)
    m.compute_dtype = torch.float32 if storage is None else torch.bfloat16
    m.train()
    # replicate the module's forward, keeping intermediates
    x = torch.tensor(x_np, device='cuda')
    dt = m._act_dtype(); nl = 4; L = m.layer_lengths(12); Lf = L[-1]
    pk = m._packed()
    supports = list(m.supports) + [ops.AdaptiveAdjacency.apply(m.nodevec1, m.nodevec2)]
    print('storage', storage, 'adp rel', rel(supports[-1], tr.adp))
    u = ops.StartConv.apply(x, m.start_conv.weight, m.start_conv.bias, L[0], dt == torch.bfloat16)
    stats = None
    for i in range(nl):
        bn_prev = m.bn[i - 1] if i > 0 else None
        meta = dict(training=True, momentum=0.1, eps=1e-5, Lf=Lf, taps=2, dilation=m.dilations[i], order=2, has_gconv=True, dropout_p=0.0, seed=0, offset=i)
        u, stats, zl = ops.WaveNetLayer.apply(u, stats, None if bn_prev is None else bn_prev.weight, None if bn_prev is None else bn_prev.bias,
            None if bn_prev is None else bn_prev.running_mean, None if bn_prev is None else bn_prev.running_var,
            pk['w_fg'][i], pk['b_fg'][i], pk['w_mlp'][i], pk['b_mlp'][i], None, None, None, meta, *supports)
        u_ref = tr.u[i].permute(0, 3, 2, 1)          # NCHW -> N,L,V,C
        z_ref = tr.z[i][..., -Lf:].permute(0, 3, 2, 1)
        if storage is not None:
            u_ref = u_ref.to(storage).double()
        d = (u.double().cpu() - u_ref)
        frac = (d.abs() > 1e-6 * u_ref.abs().clamp_min(1e-3)).double().mean().item()
        print(f' layer {i}: u rel {rel(u, u_ref):.2e} (elements differing {frac:.3%})  z_last rel {rel(zl, z_ref):.2e}')
