"""Warm per-kernel durations of the captured training step (torch.profiler over CUDA-graph replays): sum of kernel
time vs wall step time = how much of the step is launch gaps / tails.   usage: gpu_step_profile.py [c2|c3|c4|c5] [replays]"""
import os, sys, collections
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
import bench
key = sys.argv[1] if len(sys.argv) > 1 else 'c2'
R = int(sys.argv[2]) if len(sys.argv) > 2 else 5
bench.select_config(key)
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
tr = bench.Trainer(dev, 1, 0, n_batches=2)
for i in range(3):
    tr.step_resident(i)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(R):
        tr.step_resident(i)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    agg[e.name[:90]][0] += 1; agg[e.name[:90]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
print(f'{key}: {R} replays: kernel time sum {tot/R:.0f} us/step, span {(t1-t0)/R:.0f} us/step, {len(ev)/R:.0f} device activities/step')
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f'{t/R:9.1f} us {c/R:6.1f}x  {k}')

if os.environ.get('PER_INSTANCE'):        # durations of every launch of the main kernel families, in launch order (one replay)
    fams = ['gcn_bwd_t', 'gcn_fwd_t', 'EpiGateBwdTC', 'EpiGateTC', 'bn_bwd']
    ev1 = sorted(ev, key=lambda e: e.time_range.start)
    n1 = len(ev1) // R
    for f in fams:
        print(f, ' '.join(f'{e.device_time:.1f}' for e in ev1[:n1] if f in e.name))
