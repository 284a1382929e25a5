"""FlatAdam (one-launch Adam over a flat parameter buffer; at world size > 1 the gradient all-reduce over NVSwitch peer
memory happens inside the same kernel - csrc/peer.cu) against torch.optim.Adam (+ an NCCL all-reduce), the update the
reference's train step applies (lit.py:59-61)."""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _model(dev):
    from multimodal_outage_b200 import gwnet
    torch.manual_seed(0)
    return gwnet(dev, num_nodes=67, dropout=0.0, supports=[torch.eye(67), torch.eye(67)], in_dim=2, out_dim=12,
                 kernel_size=2, blocks=2, layers=2, skip_channels=64, end_channels=128)


def _fake_backward(params, step, rank):
    """deterministic stand-in gradients (same on the reference path): grad = sin(p * k + step) * (rank + 1)"""
    for i, p in enumerate(params):
        p.grad = torch.sin(p.detach() * (1.0 + 0.1 * i) + step) * (rank + 1.0)


def test_flat_adam_matches_torch_adam_single_gpu():
    from multimodal_outage_b200.flat_adam import FlatAdam
    from multimodal_outage_b200.ddp import plan_buckets
    m1, m2 = _model('cuda'), _model('cuda')
    names = [n for b in plan_buckets(m1) for n in b]
    p1, p2 = dict(m1.named_parameters()), dict(m2.named_parameters())
    ref = torch.optim.Adam([p2[n] for n in names], lr=1e-3)
    opt = FlatAdam(m1, lr=1e-3)
    assert opt.n == sum(p1[n].numel() for n in names)
    for n in names:                                               # re-homing the parameters changed no value
        assert torch.equal(p1[n], p2[n])
    for step in range(5):
        _fake_backward([p1[n] for n in names], step, 0)
        _fake_backward([p2[n] for n in names], step, 0)
        opt.step(); ref.step()
    assert opt.steps_done == 5
    for n in names:
        err = (p1[n] - p2[n]).abs().max().item()
        assert err <= 2e-6 * max(1.0, p2[n].abs().max().item()), (n, err)
    untouched = [n for n in p1 if n not in names]
    for n in untouched:                                           # parameters the block never uses are left alone
        assert torch.equal(p1[n], p2[n])
    # a real training step through the module: forward sees the re-homed parameters, backward fills .grad, step moves them
    x = torch.randn(4, 2, 67, 12, device='cuda')
    before = m1.start_conv.weight.detach().clone()
    opt.zero_grad()
    m1(x).square().mean().backward()
    opt.step()
    assert not torch.equal(before, m1.start_conv.weight)
    assert m1.start_conv.weight.data_ptr() >= opt.flat_p.data_ptr()
    opt.close()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def fused_exchange_check(rank, world, steps=4):
    """Runs on every rank (under torchrun or mp.spawn, process group initialised): the fused all-reduce + Adam kernel
    against NCCL all_reduce(AVG) + torch.optim.Adam on identical replicas; also checks the replicas stay bit-identical."""
    import torch.distributed as dist
    from multimodal_outage_b200.flat_adam import FlatAdam
    from multimodal_outage_b200.ddp import plan_buckets
    dev = torch.device('cuda', torch.cuda.current_device())
    m1, m2 = _model(dev), _model(dev)
    names = [n for b in plan_buckets(m1) for n in b]
    p1, p2 = dict(m1.named_parameters()), dict(m2.named_parameters())
    ref = torch.optim.Adam([p2[n] for n in names], lr=1e-3)
    opt = FlatAdam(m1, lr=1e-3)
    for step in range(steps):
        _fake_backward([p1[n] for n in names], step, rank)
        _fake_backward([p2[n] for n in names], step, rank)
        for n in names:
            dist.all_reduce(p2[n].grad, op=dist.ReduceOp.AVG)
        opt.step(); ref.step()
    torch.cuda.synchronize()
    worst = 0.0
    for n in names:
        err = (p1[n] - p2[n]).abs().max().item() / max(1.0, p2[n].abs().max().item())
        worst = max(worst, err)
    flat = opt.flat_p.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    identical = all(torch.equal(gathered[0], g) for g in gathered)
    # the same through a CUDA graph (device-side step counter, flags carry the step number)
    g = torch.cuda.CUDAGraph()
    _fake_backward([p1[n] for n in names], 99, rank)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        opt.step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        opt.step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    steps_done = opt.steps_done
    dist.barrier()
    opt.close()
    return worst, identical, steps_done


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        out.put((rank, fused_exchange_check(rank, world)))
    except Exception as e:                                        # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs on one node')
def test_fused_allreduce_adam_two_gpus():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        assert isinstance(res[r], tuple), res[r]
        worst, identical, steps_done = res[r]
        assert worst <= 2e-6 and identical and steps_done == 4 + 1 + 3, res[r]      # eager steps + the pre-capture step + 3 replays


if __name__ == '__main__':            # torchrun --nproc-per-node N tests/test_gpu_flat_adam.py
    import torch.distributed as dist
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    print(rank, fused_exchange_check(rank, world), flush=True)
    dist.destroy_process_group()
