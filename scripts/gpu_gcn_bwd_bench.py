"""Microbenchmark of the fused diffusion backward: node-major kernel vs transposed (T-form) kernel, config-2 layer shapes
(graph-captured, rotating buffer sets larger than L2)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
import bench
from multimodal_outage_b200 import ops, _lib
lib = _lib.lib(); dev = 'cuda'; bf = torch.bfloat16
st = lambda: torch.cuda.current_stream().cuda_stream
V, N, R = 67, 512, 4
sups = [torch.softmax(torch.randn(V, V, device=dev), dim=1) for _ in range(3)]
mats = ops.hop_mats(sups)
w_mlp = torch.randn(224, 32, device=dev) / 15
ws_w = torch.empty(131072, device=dev, dtype=torch.uint8)
for Lout in (12, 6, 1):
    dus = [torch.randn(N, Lout, V, 32, device=dev).to(bf) for _ in range(R)]
    As = [torch.tanh(torch.randn(N, Lout, V, 32, device=dev)).to(bf) for _ in range(R)]
    Bs = [torch.sigmoid(torch.randn(N, Lout, V, 32, device=dev)).to(bf) for _ in range(R)]
    dzl = torch.randn(N, 1, V, 32, device=dev).to(bf)
    dfg = [torch.empty(N, Lout, V, 64, device=dev, dtype=bf) for _ in range(R)]
    dw, db = torch.zeros(224, 32, device=dev), torch.zeros(32, device=dev)
    dA, dQ = torch.zeros(V, V, device=dev), torch.zeros(V, V, device=dev)
    for drop in (0.3, 0.0):
        for sa in (2, -1):
            res = {}
            for name in ('node', 'T'):
                def mk(i):
                    def f():
                        if name == 'T':
                            _lib.check(lib.gwn_gcn_bwd_t(dus[i].data_ptr(), As[i].data_ptr(), Bs[i].data_ptr(), dzl.data_ptr(),
                                mats.data_ptr(), 3, w_mlp.data_ptr(), drop, 42, i, sa, dfg[i].data_ptr(), dw.data_ptr(),
                                db.data_ptr(), dA.data_ptr(), dQ.data_ptr(), N, V, Lout, 1, st()), 'bwd_t')
                        else:
                            _lib.check(lib.gwn_gcn_bwd(dus[i].data_ptr(), As[i].data_ptr(), Bs[i].data_ptr(), dzl.data_ptr(),
                                mats.data_ptr(), 3, w_mlp.data_ptr(), ws_w.data_ptr(), drop, 42, i, sa, dfg[i].data_ptr(),
                                dw.data_ptr(), db.data_ptr(), dA.data_ptr(), N, V, Lout, 1, st()), 'bwd')
                    return f
                res[name] = bench.graph_time([mk(i) for i in range(R)]) * 1e3
            P = N * Lout * V
            gbs = 5 * 64 * P / (res['T'] * 1e-6) / 1e9
            print(f'Lout={Lout} drop={drop} sa={sa}: node-major {res["node"]:.1f} us  T-form {res["T"]:.1f} us '
                  f'({gbs:.0f} GB/s algorithmic, incl. 2 memsets)', flush=True)
