"""Per-role clock64 timeline of CTA 0 of the fused gcn forward kernel (debug aid)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
dev = 'cuda'
trace = torch.zeros(64 * 8, device=dev, dtype=torch.int64)
os.environ['GWN_GCN_TRACE'] = str(trace.data_ptr())
from multimodal_outage_b200 import ops, _lib
lib = _lib.lib(); bf = torch.bfloat16
st = lambda: torch.cuda.current_stream().cuda_stream
V, N = 67, 512
Lout = int(sys.argv[1]) if len(sys.argv) > 1 else 12
Lin = Lout + 1
sups = [torch.softmax(torch.randn(V, V, device=dev), dim=1) for _ in range(3)]
mats = ops.hop_mats(sups)
z = torch.randn(N, Lout, V, 32, device=dev).to(bf); up = torch.randn(N, Lin, V, 32, device=dev).to(bf)
u = torch.empty(N, Lout, V, 32, device=dev, dtype=bf)
w_mlp = torch.randn(224, 32, device=dev) / 15; b_mlp = torch.zeros(32, device=dev)
scale, shift = torch.ones(32, device=dev), torch.zeros(32, device=dev)
ws_w = torch.empty(32768, device=dev, dtype=torch.uint8); stats = torch.zeros(64, device=dev, dtype=torch.float64)
for it in range(2):
    _lib.check(lib.gwn_gcn_fwd(z.data_ptr(), up.data_ptr(), scale.data_ptr(), shift.data_ptr(), mats.data_ptr(), 3,
        w_mlp.data_ptr(), b_mlp.data_ptr(), ws_w.data_ptr(), 0.3, 42, 0, u.data_ptr(), stats.data_ptr(), N, V, Lin, Lout, st()), 'gcn')
    torch.cuda.synchronize()
t = trace.cpu().reshape(64, 8)
t0 = t[0, 1].item()
print('marks relative to kernel begin: after_prologue, kernel_end, first mma_U_start:', [t[63, i].item() - t[63, 0].item() for i in (1, 2)], t0 - t[63, 0].item())
names = ['prod_issued', 'mma_U_start', 'mma_U_issued', 'mma_hops_issued', 'stage_start', 'stage_end', 'epi_start', 'epi_end']
if os.environ.get('GWN_GCN_T', '1') != '0':
    names = ['mma_grp_start', 'mma_g1_issued', 'mma_got_a_full', 'mma_g2_issued', 'stage_start', 'stage_end', 'epi_got_d', 'epi_end']
    t0 = t[0, 0].item()
print('slab ' + ' '.join(f'{n:>15s}' for n in names))
for k in range(0, 42):
    if t[k, 1].item() == 0:
        break
    print(f'{k:4d} ' + ' '.join(f'{(t[k, j].item() - t0) if t[k, j].item() else 0:15d}' for j in range(8)))
