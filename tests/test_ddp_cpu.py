"""Data-parallel host logic on CPU (no GPU needed): the static bucket plan and a world-size-2 `gloo` run of
BucketedGradAllReduce (batch data-parallel, ONE gradient exchange per step, per-replica BatchNorm)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _model():
    from multimodal_outage_b200 import gwnet
    torch.manual_seed(0)
    return gwnet('cpu', num_nodes=11, dropout=0.0, supports=[torch.eye(11), torch.eye(11)], in_dim=2, out_dim=3,
                 kernel_size=2, blocks=2, layers=2, skip_channels=64, end_channels=64)


def test_bucket_plan_is_static_and_skips_never_used_parameters():
    from multimodal_outage_b200.ddp import plan_buckets
    m = _model()
    plan = plan_buckets(m)
    assert len(plan) == 2
    head, rest = plan
    nl = m.blocks * m.layers
    assert all(n.startswith(('end_conv_1.', 'end_conv_2.', 'skip_convs.')) for n in head)
    assert len(head) == 4 + 2 * nl
    flat = head + rest
    assert len(set(flat)) == len(flat)
    # parameters the reference never gives a gradient (graph_wavenet.py:245, :250-252) are in no bucket
    assert not any(n.startswith('residual_convs.') for n in flat)
    assert not any(n.startswith((f'bn.{nl - 1}.', f'gconv.{nl - 1}.')) for n in flat)
    for must in ('nodevec1', 'nodevec2', 'start_conv.weight', 'filter_convs.0.weight', 'gconv.0.mlp.mlp.bias', 'bn.0.weight'):
        assert must in rest
    assert plan == plan_buckets(_model())          # same plan on every rank: depends on names only


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, overlap, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from multimodal_outage_b200.ddp import BucketedGradAllReduce, plan_buckets
        m = _model()
        sync = BucketedGradAllReduce(m, overlap=overlap)
        used = set(n for b in plan_buckets(m) for n in b)
        params = dict(m.named_parameters())
        for step in range(2):                                  # two steps: buckets re-arm after finish()
            for p in m.parameters():
                p.grad = None
            # a stand-in backward (the CUDA block cannot run here): grad of p = (rank + 1 + step) * p
            loss = sum(((rank + 1.0 + step) * 0.5) * (params[n] ** 2).sum() for n in sorted(used))
            loss.backward()
            sync.finish()
            mean_scale = sum(r + 1.0 + step for r in range(world)) / world
            for n, p in params.items():
                if n in used:
                    assert p.grad is not None
                    assert torch.allclose(p.grad, mean_scale * p.detach(), rtol=1e-6, atol=1e-7), n
                else:
                    assert p.grad is None, n
        assert sync.grad_bytes() == 4 * sum(params[n].numel() for n in used)
        # CUDA-graph split used by bench.py at N > 1: hooks off, pack() (device copies only) then reduce() (the exchange)
        sync.remove()
        for p in m.parameters():
            p.grad = None
        sum((rank + 3.0) * 0.5 * (params[n] ** 2).sum() for n in sorted(used)).backward()
        sync.pack()
        sync.reduce()
        mean_scale = sum(r + 3.0 for r in range(world)) / world
        for n in used:
            assert torch.allclose(params[n].grad, mean_scale * params[n].detach(), rtol=1e-6, atol=1e-7), n
        out.put((rank, 'ok'))
    except Exception as e:                                     # surface the failure in the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _run(overlap):
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res


def test_gloo_world_size_2_bucketed_allreduce_overlapped_with_backward():
    _run(overlap=True)


def test_gloo_world_size_2_bucketed_allreduce_at_finish():
    _run(overlap=False)
