"""clock64 timeline of CTA 0 of the T-form fused backward (needs a GWN_TRACE=1 build: `GWN_TRACE=1 python -m
multimodal_outage_b200.build --force`).  Rows = items (8 per group of four slabs at V=67 with the adaptive gradient)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
dev = 'cuda'
trace = torch.zeros(256 * 16, device=dev, dtype=torch.int64)
os.environ['GWN_GCN_TRACE'] = str(trace.data_ptr())
from multimodal_outage_b200 import ops, _lib
lib = _lib.lib(); bf = torch.bfloat16
V, N, Lout, sa = 67, 512, 12, int(os.environ.get('SA', '2'))
sups = [torch.softmax(torch.randn(V, V, device=dev), dim=1) for _ in range(3)]
mats = ops.hop_mats(sups)
w_mlp = torch.randn(224, 32, device=dev) / 15
du = torch.randn(N, Lout, V, 32, device=dev).to(bf)
a = torch.tanh(torch.randn(N, Lout, V, 32, device=dev)).to(bf)
b = torch.sigmoid(torch.randn(N, Lout, V, 32, device=dev)).to(bf)
dzl = torch.randn(N, 1, V, 32, device=dev).to(bf)
dfg = torch.empty(N, Lout, V, 64, device=dev, dtype=bf)
dw, db = torch.zeros(224, 32, device=dev), torch.zeros(32, device=dev)
dA, dQ = torch.zeros(V, V, device=dev), torch.zeros(V, V, device=dev)
for _ in range(3):
    trace.zero_()
    _lib.check(lib.gwn_gcn_bwd_t(du.data_ptr(), a.data_ptr(), b.data_ptr(), dzl.data_ptr(), mats.data_ptr(), 3,
                                 w_mlp.data_ptr(), 0.3, 42, 0, sa, dfg.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                 dA.data_ptr(), dQ.data_ptr(), N, V, Lout, 1, torch.cuda.current_stream().cuda_stream), 'bwd_t')
    torch.cuda.synchronize()
t = trace.cpu().reshape(256, 16)
t0 = int(t[0, 0])
NI = 8 if sa >= 0 else 6
print('item | S1: enter waited issued | S2: enter waited dz-waited issued | stage: t_full s_empty stored | 10..13')
for i in range(min(250, NI * 8)):
    row = [int(x) - t0 if int(x) else -1 for x in t[i]]
    print(f'{i:3d} g{i // NI} i{i % NI} | ' + ' '.join(f'{x:7d}' for x in row[:3]) + ' | ' + ' '.join(f'{x:7d}' for x in row[3:7]) +
          ' | ' + ' '.join(f'{x:7d}' for x in row[7:10]) + ' | ' + ' '.join(f'{x:7d}' for x in row[10:14]))
