"""Gated conv forward, training form, config-2 layer 0 - a few plain launches (ncu target: -k regex:pos_gemm_tc_kernel -s 2 -c 1)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
from multimodal_outage_b200 import ops
dev = 'cuda'; bf = torch.bfloat16
V, N, Lin = 67, 512, 13
sups = [torch.softmax(torch.randn(V, V, device=dev), dim=1) for _ in range(3)]
mats = ops.hop_mats(sups)
ups = [torch.randn(N, Lin, V, 32, device=dev).to(bf) for _ in range(4)]
w_fg = torch.randn(64, 64, device=dev) / 8; b_fg = torch.zeros(64, device=dev)
for i in range(4):
    ops.layer_fwd(ups[i], None, None, w_fg, b_fg, None, None, [], None, None, mats, 1, 2, 1, 2, True, False, 0.0, 0, 0)
torch.cuda.synchronize()
print('ok')
