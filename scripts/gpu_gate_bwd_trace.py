"""clock64 timeline of CTA 0 of pos_gemm_tc_kernel<EpiGateBwdTC> (gate data-gradient GEMM, config-2 layer 0) via a
layer backward: the data-gradient GEMM is the last pos_gemm_tc launch of the backward, so it owns the trace buffer."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
dev = 'cuda'
trace = torch.zeros(48 * 16, device=dev, dtype=torch.int64)
os.environ['GWN_PG_TRACE'] = str(trace.data_ptr())
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
for it in range(2):
    u, stats, zl = ops.WaveNetLayer.apply(u_prev, None, None, None, None, None, w_fg, b_fg, w_mlp, b_mlp, None, None, mats, meta, *sups)
    torch.cuda.synchronize()
    trace.zero_()
    torch.autograd.backward([u, zl], [torch.randn_like(u), torch.randn_like(zl)])
    torch.cuda.synchronize()
t = trace.cpu().reshape(48, 16)
t0 = t[0, 0].item()
names = ['prod_got_empty', 'prod_issued', 'mma_got_tempty', 'mma_got_full', 'mma_issued', 'epi_got_tfull', 'epi_done', 'epi_ldtm', '-', '-', 'epi_computed']
print('tile ' + ' '.join(f'{n:>14s}' for n in names))
for k in range(0, 30):
    if t[k, 0].item() == 0:
        break
    print(f'{k:4d} ' + ' '.join(f'{(t[k, j].item() - t0) if t[k, j].item() else 0:14d}' for j in range(11)))

m = t[47]
print('kernel phases (cycles from entry): prologue set-up', m[1].item() - m[0].item(), ' pdl wait', m[2].item() - m[1].item(), ' image build', m[3].item() - m[2].item(),
      ' first producer event', t0 - m[0].item(), ' all tiles done', m[4].item() - m[0].item())
