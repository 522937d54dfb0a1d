"""Shared implementation of the drop-in ``MARN`` modules of HybridRNN_AT / HybridRNN_ATV.

Constructor and ``forward(x)`` signatures, parameter names/shapes/registration order and the
RNG consumption order of the default initialisation follow the reference
(model/HybridRNN_ATV.py:40-82, model/HybridRNN_AT.py:40-79); the time loop
(HybridRNN_ATV.py:117-143) is replaced by one fused CUDA recurrence (``recurrence.py``).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .encoder import EncoderLayer
from .mm3 import linear3, linear_cat, _LinearBlock, _rows
from .recurrence import mab_recurrence, mab_prepack
from .streams import state_without_streams

_MODS = ("l", "a", "v")


class LSTHM(nn.Module):
    """Parameter container of one LSTHM cell (model/HybridRNN_ATV.py:12-19).  Its arithmetic
    (lines 21-37) runs inside the fused recurrence kernel when the owner network's ``forward`` is called; ``forward`` here
    is the reference's single-step signature, kept for callers that drive a cell by hand (plain tensor expressions — the
    network never calls it)."""

    def __init__(self, cell_size, in_size, hybrid_in_size):
        super().__init__()
        self.cell_size, self.in_size = cell_size, in_size
        self.W = nn.Linear(in_size, 4 * cell_size)
        self.U = nn.Linear(cell_size, 4 * cell_size)
        self.V = nn.Linear(hybrid_in_size, 4 * cell_size)

    def gate_input(self, x: torch.Tensor) -> torch.Tensor:
        """W x + bW + bU + bV for all steps at once (the hoisted, time-parallel part of line 23-27)."""
        return linear3(x, self.W.weight, self.W.bias + self.U.bias + self.V.bias)

    def forward(self, x, ctm, htm, ztm):
        """One step, HybridRNN_ATV.py:21-37: gates f|i|o|g of W x + U h + V z -> (c_t, h_t)."""
        s = self.W(x) + self.U(htm) + self.V(ztm)
        d = self.cell_size
        f, i, o = torch.sigmoid(s[:, :d]), torch.sigmoid(s[:, d:2 * d]), torch.sigmoid(s[:, 2 * d:3 * d])
        c = f * ctm + i * torch.tanh(s[:, 3 * d:])
        return c, torch.tanh(c) * o


class _EncoderBranches(torch.autograd.Function):
    """The per-modality branches of ``MabNet.forward`` — encoder layer, then the LSTHM cell's hoisted gate projection
    ``W_m x + bW + bU + bV`` written into its column block of gx[T,N,4D] — as ONE autograd node that forks to side streams and
    joins again, in the forward and in the backward.

    Each branch builds its own autograd graph under ``enable_grad`` on its stream and is differentiated by a nested
    ``torch.autograd.backward`` on that same stream, which accumulates the encoder's parameter gradients directly (their
    post-accumulate hooks fire there; ``ddp.GradAllReducer`` leaves an event for the stream that packs the bucket).  The outer
    graph only sees (x, anchor) -> (y_l, y_a, ...): no tensor ever crosses streams THROUGH autograd, so the caching allocator
    never sees a ``record_stream`` (autograd records one for every gradient it hands across streams).  That matters: deferred
    frees make the allocator grow at timing-dependent moments — measured with plain multi-stream autograd: occasional steps of
    22 ms and 103 ms (cudaMalloc inside the step) among 14.6 ms ones.  With the fork/join below every reuse is ordered by the
    streams themselves: a side stream's pool is re-used only by that stream, whose next work (this step's backward, the next
    step's forward) starts with ``wait_stream(caller)``; the caller's pool is re-used behind the joins."""

    @staticmethod
    def forward(ctx, net, build, x, anchor):
        cur = torch.cuda.current_stream(x.device)
        side = net._streams(x.device)
        M = len(net._mods)
        inner = [None] * M
        offs = [sum(net._d_in[:i]) for i in range(M)]
        goffs = [4 * sum(net._dh[:i]) for i in range(M)]
        T, N = x.shape[0], x.shape[1]
        gx = torch.empty(T * N, 4 * net.total_h_dim, device=x.device, dtype=torch.float32)     # caller's pool, before the fork
        xd = x.detach()
        # widest branch first (the host is only a few ms ahead of the device: the longest branch must not be queued last;
        # measured 14.69 vs 14.89 ms per step); the narrowest, issued last, stays on the caller's stream
        order = sorted(range(M), key=lambda i: -net._d_in[i])
        for k, i in enumerate(order):
            m, d, o = net._mods[i], net._d_in[i], offs[i]
            s = cur if k == M - 1 else side[k]
            if s != cur:
                s.wait_stream(cur)
            with torch.cuda.stream(s), torch.set_grad_enabled(build):
                xm = xd[:, :, o:o + d].permute(1, 0, 2)
                if build and x.requires_grad:
                    xm.requires_grad_(True)                       # a leaf of the branch graph: collects dL/dx of this slice
                y, _ = getattr(net, f"encoder_{m}")(xm)
                y = y.permute(1, 0, 2)                            # [T,N,d]: the rows of the time-major storage
                cell, box = getattr(net, f"lsthm_{m}"), {}
                h = _LinearBlock.apply(y.reshape(T * N, d), cell.W.weight, cell.W.bias + cell.U.bias + cell.V.bias,
                                       gx[:, goffs[i]:goffs[i] + 4 * net._dh[i]], box)
            inner[i] = (xm, h, s, box)
        for s in side:
            cur.wait_stream(s)
        ctx.order, ctx.goffs = order, goffs
        ctx.net, ctx.inner, ctx.x_grad = net, (inner if build else None), bool(build and x.requires_grad)
        ctx.x_width = x.shape[2]
        return gx.view(T, N, -1)

    @staticmethod
    def backward(ctx, dgx):
        if ctx.inner is None:
            raise RuntimeError("lsthm_b200: the encoder branches were already differentiated (retain_graph is not supported here)")
        inner, ctx.inner = ctx.inner, None
        cur = torch.cuda.current_stream(dgx.device)
        side = ctx.net._streams(dgx.device)
        dg = _rows(dgx.reshape(-1, dgx.shape[-1]))
        for i in ctx.order:                                       # each branch on the stream its forward ran on
            xm, h, s, box = inner[i]
            if not h.requires_grad:
                continue
            box["dy"] = dg[:, ctx.goffs[i]:ctx.goffs[i] + 4 * ctx.net._dh[i]]
            if s != cur:
                s.wait_stream(cur)
            with torch.cuda.stream(s):
                torch.autograd.backward(h, h.new_empty(0))
        for s in side:
            cur.wait_stream(s)
        dx = None
        if ctx.x_grad:
            parts = [(t[0].grad if t[0].grad is not None else torch.zeros_like(t[0])).permute(1, 0, 2) for t in inner]
            used = sum(q.shape[2] for q in parts)
            if used < ctx.x_width:                                # columns of x beyond the modalities' slices are never read
                parts.append(parts[0].new_zeros(parts[0].shape[0], parts[0].shape[1], ctx.x_width - used))
            dx = torch.cat(parts, dim=2)
        return None, None, dx, None


class MabNet(nn.Module):
    def __getstate__(self):
        return state_without_streams(self)          # cached CUDA streams are not part of the module's state

    def __init__(self, d_in: Sequence[int], dh: Sequence[int], reduce: Sequence[int], output_dim: int):
        super().__init__()
        self._mods = _MODS[:len(d_in)]
        self.num_atts = 4
        self.total_h_dim = sum(dh)
        self.total_reduce_dim = sum(reduce)
        self._d_in, self._dh, self._rd = tuple(d_in), tuple(dh), tuple(reduce)
        h_out, map_h = 32, 64
        self._map_h = map_h
        D = self.total_h_dim
        # --- same construction order as the reference so a fixed seed gives the same weights ---
        for m, d, h in zip(self._mods, d_in, dh):
            setattr(self, f"lsthm_{m}", LSTHM(h, d, D))
        self.att = nn.Sequential(nn.Linear(D, self.num_atts * D))
        for m, h, r in zip(self._mods, dh, reduce):
            setattr(self, f"reduce_dim_nn_{m}", nn.Sequential(nn.Linear(self.num_atts * h, r)))
        self.fc = nn.Sequential(nn.Linear(self.total_reduce_dim, map_h), nn.ReLU(), nn.Dropout(0.3),
                                nn.Linear(map_h, D))
        self.nn_out = nn.Sequential(nn.Linear(2 * D, h_out), nn.ReLU(), nn.Dropout(0.0),
                                    nn.Linear(h_out, output_dim), nn.Softmax(dim=-1))
        for m, d in zip(self._mods, d_in):
            setattr(self, f"encoder_{m}", EncoderLayer(d, 50, 8, 40, 40))
        self.rows_per_cta = 0            # 0 = let the library pick the tile height
        self.concurrent_encoders = True  # issue the per-modality encoders on their own CUDA streams (see encode())
        self._side_streams = {}
        self.fc_mask_override: Optional[torch.Tensor] = None   # test hook: dropout mask tape [T,N,map_h]

    # -- pieces --------------------------------------------------------------------------------
    def recurrence_weights(self):
        cells = [getattr(self, f"lsthm_{m}") for m in self._mods]
        red = [getattr(self, f"reduce_dim_nn_{m}")[0] for m in self._mods]
        return ([c.U.weight for c in cells] + [c.V.weight for c in cells] + [self.att[0].weight, self.att[0].bias]
                + [r.weight for r in red] + [r.bias for r in red]
                + [self.fc[0].weight, self.fc[0].bias, self.fc[3].weight, self.fc[3].bias])

    def encode(self, x: torch.Tensor):
        """Per-modality slices through their EncoderLayer (HybridRNN_ATV.py:86-96); returns [T,N,d_m] each (single-stream
        schedule; ``forward`` on a CUDA device runs the branches concurrently, see ``_branches``)."""
        xs, o = [], 0
        for m, d in zip(self._mods, self._d_in):
            y, _ = getattr(self, f"encoder_{m}")(x[:, :, o:o + d].permute(1, 0, 2))
            xs.append(y.permute(1, 0, 2))
            o += d
        return xs

    def _branches(self, x: torch.Tensor) -> torch.Tensor:
        """gx[T,N,4D] = gate_inputs(encode(x)) with each modality's encoder + gate projection issued on its own stream, forward
        and backward (``_EncoderBranches``): the narrow kernels of one branch (d = 100 row kernels, weight packs, split-K
        reduces, the partial last wave of every GEMM) fill the SMs the other branches leave idle.  The streams fork from and
        join the caller's stream, so callers see ordinary stream semantics."""
        build = torch.is_grad_enabled() and (x.requires_grad or any(
            p.requires_grad for m in self._mods for mod in (getattr(self, f"encoder_{m}"), getattr(self, f"lsthm_{m}"))
            for p in mod.parameters()))
        anchor = torch.empty(0, device=x.device, requires_grad=True) if build else None
        return _EncoderBranches.apply(self, build, x, anchor)

    def _aux_stream(self, device):
        key = ("aux", device.index if device.index is not None else torch.cuda.current_device())
        if key not in self._side_streams:
            self._side_streams[key] = torch.cuda.Stream(device=device)
        return self._side_streams[key]

    def _streams(self, device):
        key = device.index if device.index is not None else torch.cuda.current_device()
        if key not in self._side_streams:
            self._side_streams[key] = [torch.cuda.Stream(device=device) for _ in range(len(self._mods) - 1)]
        return self._side_streams[key]

    def gate_inputs(self, xs) -> torch.Tensor:
        """[W_m x_m + bW_m + bU_m + bV_m]_m for all steps, written into the column blocks of one [T,N,4D] tensor."""
        cells = [getattr(self, f"lsthm_{m}") for m in self._mods]
        return linear_cat(list(xs), [c.W.weight for c in cells], [c.W.bias + c.U.bias + c.V.bias for c in cells])

    def _fc_mask(self, T, N, device):
        if self.fc_mask_override is not None:
            return self.fc_mask_override.to(device=device, dtype=torch.float32)
        p = self.fc[2].p
        if not self.training or p == 0.0:
            return None
        keep = torch.empty(T, N, self._map_h, device=device, dtype=torch.float32).bernoulli_(1.0 - p)
        return keep.mul_(1.0 / (1.0 - p))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        T, N, _ = x.shape
        pre = None
        if x.is_cuda and self.concurrent_encoders:
            # the recurrence's packed weights depend on the parameters only: built next to the encoders, off the critical path
            pre = mab_prepack(T, N, self._dh, self._rd, self._map_h, self.recurrence_weights(), self.rows_per_cta,
                              self._aux_stream(x.device))
        if x.is_cuda and x.dtype == torch.float32 and self.concurrent_encoders and len(self._mods) > 1:
            gx = self._branches(x)
        else:
            gx = self.gate_inputs(self.encode(x))
        hz = mab_recurrence(gx, self._fc_mask(T, N, x.device), self._dh, self._rd, self._map_h,
                            self.recurrence_weights(), self.rows_per_cta, prepacked=pre)
        self.last_hz = hz
        # head nn_out = Linear - ReLU - Dropout(0) - Linear - Softmax (HybridRNN_ATV.py:68-73): both products on the own GEMM
        # (ReLU in the first one's epilogue; the 6/7-class output is padded to 8 columns inside linear3)
        y = linear3(hz.view(T * N, -1), self.nn_out[0].weight, self.nn_out[0].bias, relu=True)
        y = self.nn_out[2](y)
        y = linear3(y, self.nn_out[3].weight, self.nn_out[3].bias)
        return self.nn_out[4](y)                      # time-major [T*N, C] probabilities (line 153)
