"""Data parallelism for the pretraining step: one process per GPU, gradient mean over NCCL, no data-path collective.

The reference wraps the model in DistributedDataParallel (engines/defaults.py:22-43, train.py:212-216: 25 MB buckets,
find_unused_parameters=True).  Scenes are independent and BN statistics stay per rank (sync_bn=False), so the only
exchange is the gradient mean.  Here all parameters and all gradients live in two flat fp32 buffers (the modules hold
views), laid out in the order in which backward completes them, so that

  * the optimizer is one fused kernel over one tensor instead of ~350 small launches;
  * the gradient mean is a handful of large `all_reduce(AVG)` calls on a side stream, each issued as soon as its
    contiguous slice of the buffer is complete, i.e. overlapped with the rest of backward (what DDP's buckets do), with
    the 1/world scaling done by NCCL instead of a separate pass over the buffer.

Contract (what the reference trainer's idioms turn into):
  * build the optimizer with `flat.make_optimizer(torch.optim.SGD, ...)`: its `zero_grad()` zeroes the flat gradient
    buffer and keeps the views (a plain `optimizer.zero_grad(set_to_none=True)` would drop the master gradient while
    the per-parameter views silently keep accumulating);
  * `model.zero_grad()` / `p.grad = None` is tolerated: `flat.sync_grads()` (called by `all_reduce_mean()` and by the
    optimizer's step pre-hook) copies stray per-parameter gradients into the buffer and re-attaches the views;
  * parameters that never receive a gradient (the reference leaves their `.grad` None under
    find_unused_parameters=True, so SGD neither decays nor moves them) are detected at the first step and frozen.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import nn

_ALIGN = 4   # elements: every parameter starts on a 16-byte boundary (vectorised kernels take views of the buffer)


def _is_dist(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


class FlatParameters:
    """Re-homes every trainable parameter (and its gradient) of `module` into contiguous flat buffers.

    `order`: optional list of parameters in the order backward completes their gradients (first = earliest); the
    buffers are laid out in that order and cut into `num_chunks` contiguous slices for the overlapped all-reduce."""

    def __init__(self, module: nn.Module, order: Optional[Sequence[nn.Parameter]] = None, num_chunks: int = 4):
        params: List[nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        if order is not None:
            seen = {id(p) for p in order}
            assert len(seen) == len(order), "order lists a parameter twice"
            params = [p for p in order if p.requires_grad] + [p for p in params if id(p) not in seen]
        dev, dt = params[0].device, params[0].dtype
        assert all(p.device == dev and p.dtype == dt for p in params), "flat buffers need one device and dtype"
        offsets, off = [], 0
        for p in params:
            offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        total = off
        self.flat_param = torch.zeros(total, device=dev, dtype=dt)
        self.flat_grad = torch.zeros(total, device=dev, dtype=dt)
        self.params, self.offsets = params, offsets
        self._views = []
        for p, o in zip(params, offsets):
            n = p.numel()
            view = self.flat_param[o:o + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            gview = self.flat_grad[o:o + n].view(p.shape)
            p.grad = gview
            p._pv2_sink = (self, gview)     # kernels that can accumulate in place write here (spconv/pytorch.py)
            self._views.append(gview)
        self.module = module
        self.master = nn.Parameter(self.flat_param, requires_grad=True)
        self.master.grad = self.flat_grad
        # contiguous slices of roughly equal size, in completion order
        num_chunks = max(1, min(int(num_chunks), len(params)))
        target = total / num_chunks
        self.chunks, start, first = [], 0, 0
        for i, (p, o) in enumerate(zip(params, offsets)):
            end = o + (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            if (end - start >= target and len(self.chunks) < num_chunks - 1) or i == len(params) - 1:
                self.chunks.append((start, end, first, i + 1))   # element range, parameter index range
                start, first = end, i + 1
        self._chunk_of = {}
        for c, (_, _, a, b) in enumerate(self.chunks):
            for i in range(a, b):
                self._chunk_of[id(params[i])] = c
        self._pending = [b - a for (_, _, a, b) in self.chunks]
        self._overlap = False
        self._hooks = []
        self._comm_stream = None
        self._works = []
        self._launched = [False] * len(self.chunks)
        self._next = 0          # next slice to hand to the collective (slices go out in buffer order on every rank)
        self._aux_streams = []  # side streams that write gradients straight into the buffer
        self._frozen = None     # (indices, saved values) of parameters that never get a gradient
        self._steps = 0

    # ------------------------------------------------------------------------------------------ gradients
    def zero_grad(self) -> None:
        self.flat_grad.zero_()
        self._reset_round()

    def _reset_round(self) -> None:
        self._pending = [b - a for (_, _, a, b) in self.chunks]
        self._launched = [False] * len(self.chunks)
        self._works = []
        self._next = 0

    def sync_grads(self) -> None:
        """Make `flat_grad` hold every gradient and every `p.grad` a view of it again (after `model.zero_grad()`,
        `optimizer.zero_grad(set_to_none=True)` or an external `p.grad = ...`)."""
        if self.master.grad is not self.flat_grad:
            self.master.grad = self.flat_grad
        for p, v in zip(self.params, self._views):
            g = p.grad
            if g is v:
                continue
            if g is not None and g.data_ptr() != v.data_ptr():
                v.copy_(g)          # a fresh gradient tensor autograd created after the view was dropped
            p.grad = v

    def make_optimizer(self, opt_cls: Callable, **kwargs) -> torch.optim.Optimizer:
        """`opt_cls([master], **kwargs)` whose `zero_grad` zeroes the flat buffer (views stay attached) and whose step
        first re-attaches stray gradients and afterwards restores never-touched parameters."""
        flat = self

        class _FlatOptimizer(opt_cls):   # type: ignore[misc, valid-type]
            def zero_grad(self, set_to_none: bool = True) -> None:   # noqa: ARG002 - the buffer is zeroed, never dropped
                flat.zero_grad()

        opt = _FlatOptimizer([self.master], **kwargs)
        opt.register_step_pre_hook(lambda *_: flat._before_step())
        opt.register_step_post_hook(lambda *_: flat._after_step())
        return opt

    def note_aux_stream(self, stream) -> None:
        """A side stream is writing gradients into the buffer; the all-reduce and the optimizer wait for it."""
        if stream not in self._aux_streams:
            self._aux_streams.append(stream)

    def _join_aux(self, waiter) -> None:
        for s in self._aux_streams:
            waiter.wait_stream(s)

    def _before_step(self) -> None:
        self.sync_grads()
        if self.flat_grad.is_cuda:
            self._join_aux(torch.cuda.current_stream(self.flat_grad.device))
        self.wait_all_reduce()

    def _after_step(self) -> None:
        self._steps += 1
        if self._frozen is not None:
            for i, val in zip(*self._frozen):
                self.params[i].data.copy_(val)

    def freeze_untouched(self, names_or_params: Iterable) -> None:
        """Parameters the losses never reach (the reference: `.grad` stays None, so weight decay and momentum skip
        them): their values are restored after every optimizer step."""
        named = dict(self.module.named_parameters())
        idx = []
        for x in names_or_params:
            p = named[x] if isinstance(x, str) else x
            i = next((i for i, q in enumerate(self.params) if q is p), None)
            if i is not None:              # parameters without requires_grad are not in the buffers anyway
                idx.append(i)
        self._frozen = (idx, [self.params[i].detach().clone() for i in idx])

    # ------------------------------------------------------------------------------------------ collective
    def all_reduce_mean(self, group=None) -> None:
        """The gradient mean (a23).  Without `enable_overlap()` this is one blocking-order all-reduce of the
        whole buffer; with it, only the slices whose completion hook has not fired yet are reduced here and the
        launching stream then waits for the side stream."""
        self.sync_grads()
        if self.flat_grad.is_cuda:
            self._join_aux(torch.cuda.current_stream(self.flat_grad.device))
        if not _is_dist(group):
            return
        if not self._overlap:
            w = self._reduce(self.flat_grad, group)
            if w is not None:
                w.wait()          # stream-ordered: the launching stream waits for NCCL's
            return
        for c in range(len(self.chunks)):
            if not self._launched[c]:
                self._launch_chunk(c, group)
        self.wait_all_reduce()

    def _reduce(self, t: torch.Tensor, group=None):
        world = dist.get_world_size(group)
        if dist.get_backend(group) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=True)   # 1/world inside NCCL
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True)
        w.wait()
        t.mul_(1.0 / world)
        return None

    def enable_overlap(self, group=None) -> None:
        """Issue each slice's all-reduce on a side stream from a gradient hook, as soon as every parameter of the slice
        has its gradient (completion order = buffer order, so early slices go out while backward is still running)."""
        if self._overlap or not _is_dist(group):   # single process: nothing to overlap, no per-parameter hooks
            return
        self._overlap, self._group = True, group
        if self.flat_grad.is_cuda:
            self._comm_stream = torch.cuda.Stream(device=self.flat_grad.device)
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p: nn.Parameter) -> None:
        self.mark_ready(p)

    def mark_ready(self, p: nn.Parameter) -> None:
        """`p`'s gradient is final for this step (called by the autograd hook, or by a kernel wrapper that wrote the
        gradient straight into the buffer)."""
        c = self._chunk_of.get(id(p))
        if c is None or not self._overlap:
            return
        self._pending[c] -= 1
        # collectives must be issued in the same order on every rank: slices go out strictly in buffer order
        while self._next < len(self.chunks) and self._pending[self._next] <= 0 and _is_dist(self._group):
            if not self._launched[self._next]:
                self._launch_chunk(self._next, self._group)
            self._next += 1

    def _launch_chunk(self, c: int, group=None) -> None:
        s, e, a, b = self.chunks[c]
        for i in range(a, b):   # stray gradient tensors of this slice into the buffer first
            p, v = self.params[i], self._views[i]
            if p.grad is not None and p.grad is not v and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
        self._launched[c] = True
        view = self.flat_grad[s:e]
        if self._comm_stream is not None:
            self._comm_stream.wait_stream(torch.cuda.current_stream(self.flat_grad.device))
            self._join_aux(self._comm_stream)
            with torch.cuda.stream(self._comm_stream):
                w = self._reduce(view, group)
        else:
            w = self._reduce(view, group)
        if w is not None:
            self._works.append(w)

    def wait_all_reduce(self) -> None:
        for w in self._works:
            w.wait()
        self._works = []
        if self._comm_stream is not None and self._overlap:
            torch.cuda.current_stream(self.flat_grad.device).wait_stream(self._comm_stream)

    def optimizer_params(self) -> Iterable[nn.Parameter]:
        return [self.master]


def broadcast_parameters(flat: FlatParameters, src: int = 0, group=None) -> None:
    """Identical replicas: parameters AND buffers (BatchNorm running statistics) from `src`."""
    if _is_dist(group):
        dist.broadcast(flat.flat_param, src=src, group=group)
        for b in flat.module.buffers():
            if b.numel():
                dist.broadcast(b, src=src, group=group)
