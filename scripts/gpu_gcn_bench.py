"""Microbenchmark of the fused gcn forward kernel (graph-captured, rotating buffers)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import numpy as np, torch
import bench
from multimodal_outage_b200 import ops, _lib
lib = _lib.lib(); dev = 'cuda'; bf = torch.bfloat16
st = lambda: torch.cuda.current_stream().cuda_stream
V, N, Lin, Lout, R = 67, 512, 13, 12, 4
sups = [torch.softmax(torch.randn(V, V, device=dev), dim=1) for _ in range(3)]
mats = ops.hop_mats(sups)
zs = [torch.randn(N, Lout, V, 32, device=dev).to(bf) for _ in range(R)]
ups = [torch.randn(N, Lin, V, 32, device=dev).to(bf) for _ in range(R)]
us = [torch.empty(N, Lout, V, 32, device=dev, dtype=bf) for _ in range(R)]
w_mlp = torch.randn(224, 32, device=dev) / 15; b_mlp = torch.zeros(32, device=dev)
scale, shift = torch.ones(32, device=dev), torch.zeros(32, device=dev)
ws_w = torch.empty(32768, device=dev, dtype=torch.uint8); stats = torch.zeros(64, device=dev, dtype=torch.float64)
for drop in (0.0, 0.3):
    for n_eff in (512, 256, 64):
        def gcn(i):
            def f():
                _lib.check(lib.gwn_gcn_fwd(zs[i].data_ptr(), ups[i].data_ptr(), scale.data_ptr(), shift.data_ptr(),
                    mats.data_ptr(), 3, w_mlp.data_ptr(), b_mlp.data_ptr(), ws_w.data_ptr(), drop, 42, i,
                    us[i].data_ptr(), stats.data_ptr(), n_eff, V, Lin, Lout, st()), 'gcn')
            return f
        ms = bench.graph_time([gcn(i) for i in range(R)])
        slabs = n_eff * Lout
        print(f'drop={drop} N={n_eff} slabs={slabs}: {ms*1e3:.1f} us  ({ms*1e3/ (slabs/148):.3f} us per slab-per-SM, {ms*1e-3*1.9e9/(slabs/148):.0f} cycles/slab)', flush=True)
