"""Where do the small ATen launches of one eager training step come from?  torch.profiler with Python stacks: every
aten op that launches a kernel, with the innermost frames of this repo."""
import os, sys, collections
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import torch
import bench
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
tr = bench.Trainer(dev, 1, 0, use_graph=False)
for _ in range(2):
    tr.step_resident(0)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True,
             experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
    tr.step_resident(1)
    torch.cuda.synchronize()
agg = collections.Counter()
for e in prof.events():
    if e.device_type != torch.autograd.DeviceType.CPU or not e.name.startswith('aten::'):
        continue
    if len(e.kernels) == 0:
        continue
    if e.cpu_children and any(c.name.startswith('aten::') for c in e.cpu_children):
        continue          # count the leaf op only
    frames = [f for f in (e.stack or []) if 'multimodal_outage_b200' in f or 'bench.py' in f]
    agg[(e.name + ' ' + str(list(e.input_shapes))[:60] if hasattr(e, 'input_shapes') else e.name, ' <- '.join(f.split('/')[-1] for f in frames[:3]))] += 1
for (name, where), c in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f'{c:3d}x {name:28s} {where}')
