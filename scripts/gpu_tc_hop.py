"""First light for the tcgen05 hop kernel: every image variant vs an fp64 reference."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
from multimodal_outage_b200 import ops
torch.manual_seed(0)
for V, slabs in ((67, 8), (67, 37), (80, 5), (33, 300), (67, 6144)):
    sups = [torch.softmax(torch.randn(V, V, device='cuda'), dim=1) for _ in range(3)]
    mats = ops.hop_mats(sups)
    buf = torch.randn(slabs * V, 224, device='cuda').to(torch.bfloat16)
    x = buf[:, :32].double().reshape(slabs, V, 32)
    worst = 0.0
    for s in range(3):
        A = sups[s].double()
        refs = [torch.einsum('vw,svc->swc', A, x), torch.einsum('vw,svc->swc', A @ A, x),
                torch.einsum('wv,svc->swc', A, x), torch.einsum('wv,svc->swc', A @ A, x)]
        for variant in range(4):
            m = 4 * s + variant
            slot = 1 + (m % 6)
            ops.hop_tc(mats, 12, m, buf, 0, slot, V)
            torch.cuda.synchronize()
            y = buf[:, 32 * slot:32 * slot + 32].double().reshape(slabs, V, 32)
            err = ((y - refs[variant]).norm() / refs[variant].norm()).item()
            worst = max(worst, err)
            if err > 1e-2:
                print(f'  V={V} slabs={slabs} mat {m}: rel err {err:.3e}  (y norm {y.norm():.3e} ref {refs[variant].norm():.3e})')
    # other slots untouched? slot 0 must be intact
    assert torch.equal(buf[:, :32].double().reshape(slabs, V, 32), x)
    print(f'V={V} slabs={slabs}: worst rel err {worst:.3e}', flush=True)
# timing at the bench shape
V, slabs = 67, 6144
sups = [torch.softmax(torch.randn(V, V, device='cuda'), dim=1)]
mats = ops.hop_mats(sups)
buf = torch.randn(slabs * V, 224, device='cuda').to(torch.bfloat16)
for _ in range(3): ops.hop_tc(mats, 4, 0, buf, 0, 1, V)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.hop_tc(mats, 4, 0, buf, 0, 1, V)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f'hop_tc: {ms*1e3:.1f} us per hop, {2*slabs*32*V*V/ms/1e9:.1f} TFLOP/s algorithmic, {(slabs*V*64*2)/ms/1e6:.0f} GB/s')
