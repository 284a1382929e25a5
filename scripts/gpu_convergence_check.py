"""fp32 vs bf16 training trajectories of config 2 (same init, same batches, same dropout masks: the in-kernel mask is a
function of (seed, offset, element) in every kernel path).  Prints the loss every 20 steps for both precisions."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import numpy as np
import torch
from multimodal_outage_b200 import gwnet
from multimodal_outage_b200.supports import double_transition

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
adj = np.load(os.path.join(ROOT, 'tests', 'golden', 'adj_mx_fl.npy')).astype(np.float32)
sup = [torch.tensor(s) for s in double_transition(adj)]
steps, n = int(sys.argv[1]) if len(sys.argv) > 1 else 100, 256
g = torch.Generator().manual_seed(1)
xs = [torch.randn(n, 2, 67, 12, generator=g).cuda() for _ in range(4)]
# a learnable target: a fixed random linear map of the last input step
wt = torch.randn(12, 2, generator=g).cuda()
ys = [torch.einsum('oc,ncv->nov', wt, x[..., -1]).unsqueeze(-1).contiguous() for x in xs]
curves = {}
for dt in (torch.float32, torch.bfloat16):
    torch.manual_seed(42)
    m = gwnet('cuda', num_nodes=67, dropout=0.3, supports=sup, in_dim=2, out_dim=12, kernel_size=2, blocks=4, layers=2)
    m.compute_dtype = dt
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    out = []
    for i in range(steps):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(m(xs[i % 4]), ys[i % 4])
        loss.backward()
        opt.step()
        if i % 20 == 0 or i == steps - 1:
            out.append((i, float(loss)))
    curves[str(dt)] = out
for k, v in curves.items():
    print(k, ' '.join(f'{i}:{l:.4f}' for i, l in v))
a, b = curves['torch.float32'][-1][1], curves['torch.bfloat16'][-1][1]
print(f'final loss fp32 {a:.4f}  bf16 {b:.4f}  rel diff {abs(a - b) / a:.3%}')
