"""Warm per-kernel durations of the captured training step (torch.profiler over CUDA-graph replays): sum of kernel
time vs wall step time = how much of the step is launch gaps / tails."""
import os, sys, collections
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import numpy as np, torch
import bench
from multimodal_outage_b200 import gwnet
w = bench.WORKLOAD
dev = torch.device('cuda', 0)
torch.manual_seed(42)
model = gwnet(dev, num_nodes=w['V'], dropout=w['dropout'], supports=[torch.tensor(s) for s in bench.fl_supports()],
              in_dim=w['in_dim'], out_dim=w['out_dim'], kernel_size=w['kernel_size'], blocks=w['blocks'], layers=w['layers'])
model.compute_dtype = torch.bfloat16
model.train()
from multimodal_outage_b200.flat_adam import FlatAdam
opt = FlatAdam(model, lr=1e-3)
n = w['batch_per_gpu']
x = torch.randn(n, w['in_dim'], w['V'], w['T'], device=dev); y = torch.randn(n, w['out_dim'], w['V'], 1, device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(model(x), y)
    loss.backward(); opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s): step()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
with torch.cuda.graph(g): step()
for _ in range(5): g.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    agg[e.name[:70]][0] += 1; agg[e.name[:70]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
print(f'5 replays: kernel time sum {tot/5:.0f} us/step, span {(t1-t0)/5:.0f} us/step, {len(ev)/5:.0f} device activities/step')
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f'{t/5:8.1f} us {c/5:5.1f}x  {k}')
