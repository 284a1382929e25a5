"""clock64 timeline of CTA 0 of pos_gemm_tc_kernel<EpiGateTC> (gated conv forward, config-2 layer 0, eval form)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
dev = 'cuda'
trace = torch.zeros(48 * 16, device=dev, dtype=torch.int64)
os.environ['GWN_PG_TRACE'] = str(trace.data_ptr())
from multimodal_outage_b200 import ops
bf = torch.bfloat16
V, N, Lin = 67, 512, 13
sups = [torch.softmax(torch.randn(V, V, device=dev), dim=1) for _ in range(3)]
mats = ops.hop_mats(sups)
u_prev = torch.randn(N, Lin, V, 32, device=dev).to(bf)
w_fg = torch.randn(64, 64, device=dev) / 8; b_fg = torch.zeros(64, device=dev)
for _ in range(2):
    ops.layer_fwd(u_prev, None, None, w_fg, b_fg, None, None, [], None, None, mats, 1, 2, 1, 2, os.environ.get("GATE_TRAIN", "1") == "1", False, 0.0, 0, 0)
torch.cuda.synchronize()
t = trace.cpu().reshape(48, 16)
t0 = t[0, 0].item()
names = ['prod_got_empty', 'prod_issued', 'mma_got_tempty', 'mma_got_full', 'mma_issued', 'epi_got_tfull', 'epi_done', 'epi_ldtm', 'st_got_sfull', 'st_released', 'epi_computed']
print('tile ' + ' '.join(f'{n:>14s}' for n in names))
for k in range(0, 34):
    print(f'{k:4d} ' + ' '.join(f'{(t[k, j].item() - t0) if t[k, j].item() else 0:14d}' for j in range(11)))

m = t[47]
print('kernel phases (cycles from entry): prologue set-up', m[1].item() - m[0].item(), ' pdl wait', m[2].item() - m[1].item(), ' image build', m[3].item() - m[2].item(),
      ' first producer event', t0 - m[0].item(), ' all tiles done', m[4].item() - m[0].item())
