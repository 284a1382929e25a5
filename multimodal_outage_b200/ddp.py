"""Batch data-parallel training for the gwnet block: static flat gradient buckets all-reduced
(averaged) over torch.distributed (NCCL over NVLink/NVSwitch on B200; gloo in the CPU tests),
launched from post-accumulate-grad hooks so the first bucket's exchange overlaps the rest of backward.
The whole sequence - hooks, the two collectives on NCCL's stream, the waits, a fused Adam step - is
capturable in ONE CUDA graph together with forward and backward (bench.py at N > 1 does that).

The reference has no distributed code at all; this adds the one strategy the path
needs: identical replicas, batch sharded by rank, ONE exchange per step.  BatchNorm
statistics stay per replica, exactly as plain nn.BatchNorm2d does in the reference.

Bucket plan (static, identical on every rank):
  bucket 0  end_conv_2, end_conv_1, skip_convs.*   - their grads are complete at the very start of
            backward (the head runs first and its packed gradients are scattered by their own autograd
            node right after it, graph_wavenet._PackHead), ~2/3 of all gradient bytes
  bucket 1  everything else that receives a gradient (layers in reverse, start_conv, nodevec1/2)
Parameters that never receive a gradient (residual_convs.* when gcn_bool, the last layer's
bn/gconv - graph_wavenet.py:245,250-252) are left out on every rank.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def plan_buckets(model: torch.nn.Module) -> List[List[str]]:
    """Static two-bucket plan from parameter names only (no dry-run needed)."""
    nl = model.blocks * model.layers
    dead_prefixes = [f'bn.{nl - 1}.']
    if model.gcn_bool:
        dead_prefixes += ['residual_convs.', f'gconv.{nl - 1}.']
    head, rest = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad or any(name.startswith(d) for d in dead_prefixes):
            continue
        if name.startswith(('end_conv_1.', 'end_conv_2.', 'skip_convs.')):
            head.append(name)
        else:
            rest.append(name)
    return [b for b in (head, rest) if b]


class BucketedGradAllReduce:
    """Averages gradients across ranks through flat fp32 buckets.

    usage:  sync = BucketedGradAllReduce(model);  loss.backward();  sync.finish();  opt.step()
    After ``finish()`` every bucketed ``param.grad`` is a view into its (averaged) flat bucket.
    """

    def __init__(self, model: torch.nn.Module, process_group=None, overlap: bool = True):
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.overlap = overlap
        # NCCL averages inside the reduction; gloo (CPU tests) sums and the result is divided afterwards
        self._avg_in_collective = dist.is_initialized() and dist.get_backend(process_group) == 'nccl'
        params: Dict[str, torch.nn.Parameter] = dict(model.named_parameters())
        self.plan = plan_buckets(model)
        self.buckets = []
        self._owner: Dict[torch.nn.Parameter, int] = {}
        # the buckets are consecutive slices of ONE flat tensor: the overlapped (eager) mode reduces them one by one as they
        # complete, the CUDA-graph mode reduces the whole tensor with a single collective
        sizes = [sum(params[n].numel() for n in names) for names in self.plan]
        dev = params[self.plan[0][0]].device
        self.flat_all = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        start = 0
        for bi, names in enumerate(self.plan):
            ps = [params[n] for n in names]
            total = sizes[bi]
            flat = self.flat_all[start:start + total]
            start += total
            views, off = [], 0
            for p in ps:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
                self._owner[p] = bi
            self.buckets.append(dict(names=names, params=ps, flat=flat, views=views, pending=len(ps),
                                     work=None, launched=False))
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self._owner]

    # -- hook: a parameter's gradient for this step is final
    def _on_grad(self, p: torch.nn.Parameter):
        b = self.buckets[self._owner[p]]
        b['pending'] -= 1
        if b['pending'] == 0 and self.overlap:
            self._launch(b)

    def _launch(self, b):
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in b['params']]
        torch._foreach_copy_(b['views'], grads)
        b['launched'] = True
        if self.world > 1:
            op = dist.ReduceOp.AVG if self._avg_in_collective else dist.ReduceOp.SUM
            b['work'] = dist.all_reduce(b['flat'], op=op, group=self.group, async_op=True)

    def finish(self):
        """Wait for the exchanges, average, and point .grad at the bucket views."""
        for b in self.buckets:
            if not b['launched']:           # no overlap, or some gradient never arrived this step
                self._launch(b)
            b['launched'] = False
            if b['work'] is not None:
                b['work'].wait()
                b['work'] = None
            if self.world > 1 and not self._avg_in_collective:
                b['flat'].div_(self.world)
            for p, v in zip(b['params'], b['views']):
                p.grad = v
            b['pending'] = len(b['params'])

    # -- CUDA-graph friendly split: pack() is pure device copies (capturable with forward+backward), reduce() is the
    #    one NCCL exchange per step, issued eagerly right after the graph replay.
    def pack(self):
        """Copy this step's gradients into the flat buckets and point .grad at the bucket views (no communication)."""
        for b in self.buckets:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in b['params']]
            torch._foreach_copy_(b['views'], grads)
            for p, v in zip(b['params'], b['views']):
                p.grad = v
            b['pending'] = len(b['params'])
            b['launched'] = False

    def reduce(self):
        """All-reduce (average) the packed buckets in place."""
        if self.world > 1:
            # one collective over the concatenated buckets; NCCL averages in the reduction itself
            if dist.get_backend(self.group) == 'nccl':
                dist.all_reduce(self.flat_all, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(self.flat_all, op=dist.ReduceOp.SUM, group=self.group)
                self.flat_all.div_(float(self.world))

    def grad_bytes(self) -> int:
        return sum(b['flat'].numel() * 4 for b in self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()
