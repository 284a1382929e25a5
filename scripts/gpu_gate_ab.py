"""Gated conv forward (config-2 layer 0), eval and training form, live CUDA-event time per launch; GWN_PG_SUB=<1|2|4> forces the
macro-tile size of pos_gemm_tc (A/B).  Result on B200: 21-22 us eval (2.5-2.6 TB/s), 36-38 us training (2.9-3.0 TB/s) for every size."""
import os, sys, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
from multimodal_outage_b200 import ops
import bench
dev='cuda'; bf=torch.bfloat16
V,N,Lin=67,512,13
sups=[torch.softmax(torch.randn(V,V,device=dev),dim=1) for _ in range(3)]
mats=ops.hop_mats(sups)
ups=[torch.randn(N,Lin,V,32,device=dev).to(bf) for _ in range(4)]
w_fg=torch.randn(64,64,device=dev)/8; b_fg=torch.zeros(64,device=dev)
def gate(i, training):
    def f():
        ops.layer_fwd(ups[i], None, None, w_fg, b_fg, None, None, [], None, None, mats, 1, 2, 1, 2, training, False, 0.0, 0, 0)
    return f
for tr in (False, True):
    ms=bench.graph_time([gate(i, tr) for i in range(4)])
    byts=(N*32*V*Lin + N*32*V*(Lin-1)*(3 if tr else 1))*2.0
    print('SUB', os.environ.get('GWN_PG_SUB','auto'), 'training' if tr else 'eval', f'{ms*1e3:.1f} us', f'{byts/ms/1e6:.0f} GB/s')
