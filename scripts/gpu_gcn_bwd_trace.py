"""clock64 timeline of CTA 0's MMA thread in the fused diffusion backward kernel (config-2 layer-0 shape)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
dev = 'cuda'
trace = torch.zeros(64 * 8, device=dev, dtype=torch.int64)
os.environ['GWN_GCN_TRACE'] = str(trace.data_ptr())
from multimodal_outage_b200 import ops
bf = torch.bfloat16
V, N, dil = 67, 512, 1
Lin = int(sys.argv[1]) if len(sys.argv) > 1 else 13
sups = [torch.softmax(torch.randn(V, V, device=dev), dim=1) for _ in range(3)]
sups[-1].requires_grad_(True)
mats = ops.hop_mats([s.detach() for s in sups])
u_prev = torch.randn(N, Lin, V, 32, device=dev).to(bf).requires_grad_(True)
w_fg = (torch.randn(64, 64, device=dev) / 8).requires_grad_(True); b_fg = torch.zeros(64, device=dev, requires_grad=True)
w_mlp = (torch.randn(224, 32, device=dev) / 15).requires_grad_(True); b_mlp = torch.zeros(32, device=dev, requires_grad=True)
meta = dict(training=True, momentum=0.1, eps=1e-5, Lf=1, taps=2, dilation=dil, order=2, has_gconv=True, dropout_p=0.3, seed=1, offset=0)
u, stats, zl = ops.WaveNetLayer.apply(u_prev, None, None, None, None, None, w_fg, b_fg, w_mlp, b_mlp, None, None, mats, meta, *sups)
trace.zero_()
torch.autograd.backward([u, zl], [torch.randn_like(u), torch.randn_like(zl)])
torch.cuda.synchronize()
t = trace.cpu().reshape(64, 8)
t0 = t[0, 0].item()
print('marks relative to kernel begin: after_prologue, kernel_end, first got_in_full:', [t[63, i].item() - t[63, 0].item() for i in (1, 2)], t0 - t[63, 0].item())
names = ['got_in_full', 'got_ut_empty', 'hops_issued', 'Y1_issued', 'tail:got_us_full', 'tail:ZW_issued', 'iter_end']
print('slab ' + ' '.join(f'{n:>17s}' for n in names))
for k in range(0, 24):
    if t[k, 0].item() == 0:
        break
    print(f'{k:4d} ' + ' '.join(f'{(t[k, j].item() - t0) if t[k, j].item() else 0:17d}' for j in range(7)))
