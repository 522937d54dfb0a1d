"""Fork/join of independent branches of a module onto side CUDA streams, forward AND backward, as one autograd node.

Why a node of its own instead of plain multi-stream autograd (which PyTorch supports): autograd calls ``record_stream`` on every
tensor it hands from one stream to another; the caching allocator then defers those frees by polling events and, at
timing-dependent moments, has to ``cudaMalloc`` inside a training step (measured on the AT/ATV step: steps of 22 and 103 ms among
14.6 ms ones).  Here each branch builds its own graph under ``enable_grad`` on its stream and is differentiated by a nested
``torch.autograd.backward`` on that same stream, so nothing crosses streams THROUGH autograd; every reuse of memory is ordered by
the fork (``side.wait_stream(caller)``) and the join (``caller.wait_stream(side)``): a side stream's pool is only ever re-used by
that stream, whose next work starts with a fork; the caller's pool is re-used behind a join.  ``mab_net._EncoderBranches`` is the
specialised form for the AT/ATV model (its branches write column blocks of one matrix)."""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch


def side_streams(owner, device: torch.device, n: int) -> List["torch.cuda.Stream"]:
    """``n`` side streams per (module, device), created once."""
    cache = owner.__dict__.setdefault("_fork_streams", {})
    key = device.index if device.index is not None else torch.cuda.current_device()
    if len(cache.get(key, ())) < n:
        cache[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
    return cache[key][:n]


def state_without_streams(module) -> dict:
    """``__getstate__`` body for modules that cache CUDA streams on themselves: ``copy.deepcopy`` / ``pickle`` of the module must
    not try to copy the streams (they are re-created on first use)."""
    d = module.__dict__.copy()
    for k in ("_fork_streams", "_side_streams"):
        if k in d:
            d[k] = {}
    return d


class _ForkJoin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fns, streams, build, anchor, *xs):
        cur = torch.cuda.current_stream(xs[0].device)
        n, inner, outs = len(fns), [], []
        for k, (fn, x) in enumerate(zip(fns, xs)):
            s = cur if k == n - 1 else streams[k]              # the last branch stays on the caller's stream
            if s != cur:
                s.wait_stream(cur)
            with torch.cuda.stream(s), torch.set_grad_enabled(build):
                xi = x.detach()
                if build and x.requires_grad:
                    xi.requires_grad_(True)                     # leaf of the branch graph: collects dL/dx
                y = fn(xi)
            inner.append((xi, y, s))
            outs.append(y.detach())
        for s in streams[:n - 1]:
            cur.wait_stream(s)
        ctx.inner, ctx.streams = (inner if build else None), streams[:n - 1]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *dys):
        if ctx.inner is None:
            raise RuntimeError("lsthm_b200: these branches were already differentiated (retain_graph is not supported here)")
        inner, ctx.inner = ctx.inner, None
        dev = next(d for d in dys if d is not None).device
        cur = torch.cuda.current_stream(dev)
        for (xi, y, s), dy in zip(inner, dys):
            if dy is None or not y.requires_grad:
                continue
            if s != cur:
                s.wait_stream(cur)
            with torch.cuda.stream(s):
                torch.autograd.backward(y, dy)                   # accumulates the branch's parameter gradients on its stream
        for s in ctx.streams:
            cur.wait_stream(s)
        return (None, None, None, None, *[(xi.grad if xi.requires_grad else None) for xi, _, _ in inner])


def fork_join(owner, fns: Sequence[Callable[[torch.Tensor], torch.Tensor]], xs: Sequence[torch.Tensor]):
    """[fn_i(x_i)] with every branch but the last on its own stream (forward and backward).  Branch i is a function of x_i and
    of module parameters only.  Put the longest branch first."""
    dev = xs[0].device
    build = torch.is_grad_enabled()
    anchor = torch.empty(0, device=dev, requires_grad=True) if build else None
    return _ForkJoin.apply(list(fns), side_streams(owner, dev, len(fns) - 1), build, anchor, *xs)
