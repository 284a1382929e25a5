"""Adam over ONE flat parameter buffer, fused with the data-parallel gradient exchange (csrc/peer.cu).

The reference's train step ends with ``torch.optim.Adam(lr=1e-3).step()`` (lit.py:59-61) and has no distributed code
(SURVEY §2.2).  `FlatAdam` keeps the same update rule (amsgrad off, no weight decay) but
  * re-homes the parameters that receive gradients into one flat fp32 buffer (``p.data`` become views; moments flat too),
  * runs the whole update as ONE launch (``gwn_adam_flat``) instead of three multi-tensor launches over ~110 tensors, and
  * at world size > 1 does the gradient all-reduce INSIDE that launch (``gwn_allreduce_adam``): every rank's flat gradient
    lives in a CUDA-IPC exchange block mapped by all peers; the kernel barriers on flags in those blocks, reads all ranks'
    gradients over NVLink / NVSwitch peer memory, averages and applies Adam.  No NCCL call on the step path
    (torch.distributed is only used once, to trade the 64-byte IPC handles).

usage (same place as ``opt.step()``):
    opt = FlatAdam(model, lr=1e-3)                 # after the model is on its device, before any CUDA-graph capture
    loss.backward();  opt.step();  opt.zero_grad()

The step counter lives on the device: a captured ``opt.step()`` advances it on every graph replay.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from ._lib import GwnError, check, lib
from .ddp import plan_buckets


class _Foreign:
    """`__cuda_array_interface__` view of memory owned by libgwn (the exchange block's gradient area)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '<f4', 'data': (ptr, False), 'version': 2, 'strides': None}


class FlatAdam:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group=None):
        names = [n for b in plan_buckets(model) for n in b]          # parameters that ever receive a gradient
        params = dict(model.named_parameters())
        self.params: List[torch.nn.Parameter] = [params[n] for n in names]
        if not self.params or not self.params[0].is_cuda:
            raise GwnError('FlatAdam needs the model on a CUDA (B200) device - there is no CPU fallback')
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        dev = self.params[0].device
        self.device = dev
        self.n = sum(p.numel() for p in self.params)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        L = lib()
        with torch.cuda.device(dev):
            self.flat_p = torch.empty(self.n, device=dev, dtype=torch.float32)
            self.exp_avg = torch.zeros(self.n, device=dev, dtype=torch.float32)
            self.exp_avg_sq = torch.zeros(self.n, device=dev, dtype=torch.float32)
            self.state = torch.zeros(2, device=dev, dtype=torch.int64)         # [completed steps, CTA counter]
            # the gradient lives in an exchange block (header of flags + flat gradient) even at world size 1
            blk = C.c_void_p()
            check(L.gwn_peer_alloc(self.n * 4, C.byref(blk)), 'gwn_peer_alloc')
            self._block = blk.value
            self._hdr = int(L.gwn_peer_header_bytes())
            self._foreign = _Foreign(self._block + self._hdr, self.n)
            self.flat_g = torch.as_tensor(self._foreign, device=dev)
        off = 0
        self._g_views = []
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)
                self._g_views.append(self.flat_g[off:off + k].view_as(p))
                off += k
        self._counts = (C.c_longlong * len(self.params))(*[p.numel() for p in self.params])
        self._mapped: List[Optional[int]] = [None] * self.world
        self._blocks = None
        if self.world > 1:
            if self.world > 8:
                raise GwnError('the fused exchange handles up to 8 ranks (one NVSwitch domain)')
            handle = (C.c_ubyte * 64)()
            check(L.gwn_peer_export(self._block, handle), 'gwn_peer_export')
            gathered: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(gathered, bytes(handle), group=process_group)
            arr = (C.c_void_p * self.world)()
            with torch.cuda.device(dev):
                for r in range(self.world):
                    if r == self.rank:
                        arr[r] = self._block
                        continue
                    h = (C.c_ubyte * 64).from_buffer_copy(gathered[r])
                    out = C.c_void_p()
                    check(L.gwn_peer_open(h, C.byref(out)), 'gwn_peer_open')
                    self._mapped[r] = out.value
                    arr[r] = out.value
            self._blocks = arr
            # replicas must start identical: broadcast rank 0's parameters once
            dist.broadcast(self.flat_p, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                           group=process_group)
            dist.barrier(group=process_group)

    # ------------------------------------------------------------------ step
    def pack_grads(self):
        """Gathers this step's per-parameter gradients into the flat gradient of the exchange block (one launch:
        `gwn_gather_flat`, pointer table by value) and points ``.grad`` at the flat views."""
        n = len(self.params)
        if n <= 160:                                  # one launch, pointer table by value (csrc/peer.cu)
            srcs = (C.c_void_p * n)()
            for i, (p, v) in enumerate(zip(self.params, self._g_views)):
                g = p.grad
                if g is not None and g.data_ptr() == v.data_ptr():
                    srcs[i] = v.data_ptr()                          # already in place: the kernel skips src == dst
                elif g is None:
                    srcs[i] = None
                else:
                    if g.dtype != torch.float32 or not g.is_contiguous():
                        g = g.float().contiguous()
                        p.grad = g                                  # keep it alive until the launch has been enqueued
                    srcs[i] = g.data_ptr()
            with torch.cuda.device(self.device):
                check(lib().gwn_gather_flat(srcs, self._counts, n, self.flat_g.data_ptr(),
                                            torch.cuda.current_stream(self.device).cuda_stream), 'gwn_gather_flat')
        else:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
            torch._foreach_copy_(self._g_views, grads)
        for p, v in zip(self.params, self._g_views):
            p.grad = v

    def step(self):
        self.pack_grads()
        L = lib()
        st = torch.cuda.current_stream(self.device).cuda_stream
        b1, b2 = self.betas
        with torch.cuda.device(self.device):
            if self.world == 1:
                check(L.gwn_adam_flat(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), self.n, self.lr, b1, b2, self.eps, self.state.data_ptr(), st),
                      'gwn_adam_flat')
            else:
                check(L.gwn_allreduce_adam(self.flat_p.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.n,
                                           self.lr, b1, b2, self.eps, self.state.data_ptr(), self._blocks, self.rank, self.world,
                                           st), 'gwn_allreduce_adam')

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @property
    def steps_done(self) -> int:
        return int(self.state[0].item())

    def close(self):
        """Unmaps the peers' blocks and frees the own one (all ranks must be past their last step)."""
        L = lib()
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        for r, m in enumerate(self._mapped):
            if m is not None:
                L.gwn_peer_close(m)
                self._mapped[r] = None
        if self._block is not None:
            self.flat_g = None
            self._g_views = []
            L.gwn_peer_free(self._block)
            self._block = None
