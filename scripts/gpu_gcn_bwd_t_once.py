"""A few launches of the T-form fused diffusion backward at the config-2 layer-0 shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
from multimodal_outage_b200 import ops, _lib
lib = _lib.lib(); dev = 'cuda'; bf = torch.bfloat16
V, N, Lout, sa = 67, 512, 12, 2
torch.manual_seed(0)
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
for _ in range(4):
    _lib.check(lib.gwn_gcn_bwd_t(du.data_ptr(), a.data_ptr(), b.data_ptr(), dzl.data_ptr(), mats.data_ptr(), 3,
                                 w_mlp.data_ptr(), 0.3, 42, 0, sa, dfg.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                 dA.data_ptr(), dQ.data_ptr(), N, V, Lout, 1, torch.cuda.current_stream().cuda_stream), 'bwd_t')
torch.cuda.synchronize()
print('ok', float(dfg.float().abs().mean()))
