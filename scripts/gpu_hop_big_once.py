"""Runs the V=3100 diffusion hop GEMM a few times (ncu target for the config-3 roofline entry)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
from multimodal_outage_b200 import ops, _lib
lib = _lib.lib(); dev = 'cuda'
V, slabs = 3100, 768
A = torch.softmax(torch.randn(V, V, device=dev), dim=1)
img = ops.support_images([A])
xs = [torch.randn(slabs, V, 32, device=dev).to(torch.bfloat16) for _ in range(2)]
ys = [torch.empty_like(xs[0]) for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream
for i in range(4):
    _lib.check(lib.gwn_hop_big(img.data_ptr(), 1, 0, 0, xs[i % 2].data_ptr(), ys[i % 2].data_ptr(), None, slabs, V, st), 'hop')
torch.cuda.synchronize()
print('ok')
