"""torch.autograd glue around the CUDA recurrence kernels (C ABI: include/lsthm_b200.h).

`mab_recurrence` replaces the time loop of the reference's ``MARN.forward``
(model/HybridRNN_ATV.py:117-143, model/HybridRNN_AT.py:107-132).  Everything either side of
it — input projections, encoders, the per-step head — is ordinary PyTorch, so autograd
composes through ``MabRecurrenceFn.backward``.

Backward contract: the BPTT kernel returns the per-step adjoints; the weight gradients are
time-parallel products over all T*N rows (``adj^T @ act``) formed here in fp32.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .mm3 import mm_tn, mm_nn, mm_nt, colsum, linear_into

# Counts launches of OUR kernels (pack/fwd/bwd), for bench.py's `gpu_launches`.
launch_counter = {"pack": 0, "fwd": 0, "bwd": 0}
# bench.py sets this to {"fwd": [], "bwd": []} to collect (start, end) CUDA events recorded on the
# launching stream around each recurrence kernel; None = no events (the default).
kernel_events = None


_workspaces = {}


def _workspace(desc, device) -> torch.Tensor:
    """Exchange workspace of the group kernels (per device; forward and backward of a step run back to back on one stream)."""
    nbytes = _lib.mab_workspace_bytes(desc)
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    w = _workspaces.get(key)
    if w is None or w.numel() < nbytes:
        w = torch.empty(nbytes, device=device, dtype=torch.uint8)
        _workspaces[key] = w
    return w


# NVTX range names of the kernel launches (K1/K2 = AT/ATV recurrence fwd / BPTT, K3/K4 = speaker-state cell
# fwd / BPTT, K5 = time-parallel GEMMs (mm3.py), K6 = gradient allreduce (ddp.py))
NVTX_NAMES = {"mab_pack": "lsthm/K0_mab_pack", "mab_fwd": "lsthm/K1_mab_fwd", "mab_bwd": "lsthm/K2_mab_bwd",
              "sps_fwd": "lsthm/K3_sps_fwd", "sps_bwd": "lsthm/K4_sps_bwd"}


def _timed(kind, fn, *args):
    nvtx = torch.cuda.nvtx if args and any(isinstance(a, torch.Tensor) and a.is_cuda for a in args) else None
    if nvtx is not None:
        nvtx.range_push(NVTX_NAMES.get(getattr(fn, "__name__", ""), "lsthm/" + kind))
    try:
        if kernel_events is None or kind not in kernel_events:
            return fn(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(*args)
        e1.record()
        kernel_events[kind].append((e0, e1))
    finally:
        if nvtx is not None:
            nvtx.range_pop()


def _pack(desc, weights, M, device) -> torch.Tensor:
    """Composite chain weights + per-rank UMMA images for one launch pair (``lsthm_mab_pack``: two kernels, ~0.09 ms)."""
    U, V = weights[0:M], weights[M:2 * M]
    Watt, batt = weights[2 * M], weights[2 * M + 1]
    Wr, br = weights[2 * M + 2:3 * M + 2], weights[3 * M + 2:4 * M + 2]
    Wf1, bf1, Wf2, bf2 = weights[4 * M + 2:4 * M + 6]
    wstruct = _lib.make_weights(U, V, Watt, batt, Wr, br, Wf1, bf1, Wf2, bf2)
    packed = torch.empty(_lib.mab_pack_bytes(desc), device=device, dtype=torch.uint8)
    _timed("pack", _lib.mab_pack, desc, wstruct, packed)
    launch_counter["pack"] += 2
    return packed


def mab_prepack(T: int, N: int, dh: Sequence[int], rd: Sequence[int], map_h: int, weights: Sequence[torch.Tensor],
                rows_per_cta: int, stream: "torch.cuda.Stream"):
    """Build the packed weights of ``mab_recurrence`` on ``stream`` (forked from the current stream) — they depend on the
    parameters only, so a caller can have them built while the encoders run.  Returns ``(packed, ready_event)`` for
    ``mab_recurrence(..., prepacked=...)``.  ``packed`` lives in ``stream``'s pool and is next re-used by this function, behind
    its own ``wait_stream`` — no record_stream needed (see mab_net._EncoderBranches)."""
    weights = tuple(w.detach().contiguous() for w in weights)
    dev = weights[0].device
    desc = _lib.make_desc(T, N, tuple(dh), tuple(rd), int(map_h), 4, int(rows_per_cta))
    stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(stream):
        packed = _pack(desc, weights, len(dh), dev)
        ready = torch.cuda.Event()
        ready.record(stream)
    return packed, ready


class MabRecurrenceFn(torch.autograd.Function):
    """hz[T,N,2D] = recurrence(gx[T,N,4D]; U_m, V_m, att, reduce_m, fc).

    Argument order after (gx, drop_mask, dims): U_0..U_{M-1}, V_0.., Watt, batt, Wr_0.., br_0..,
    Wf1, bf1, Wf2, bf2  — all in nn.Linear layout, i.e. the modules' own parameter storage.

    The kernels run the chain with composite weights (include/lsthm_b200.h): per step  gates(h, u) -> cell ->
    attention -> u' = relu(W1 attended + b1).  What they no longer touch is formed here as time-parallel products over
    all T*N rows:  z = u Wf2^T + bf2  (forward),  duz = dz_head Wf2  (into the BPTT);  the weight gradients of the composed
    layers follow from three products with K = T*N (ds^T [h|u], dup^T attended, dz_head^T u) and a few tiny matrix products.
    """

    @staticmethod
    def forward(ctx, gx: torch.Tensor, drop_mask: Optional[torch.Tensor], dims: Tuple, *weights: torch.Tensor):
        dh, rd, map_h, rows_per_cta = dims[:4]
        prepacked = dims[4] if len(dims) > 4 else None
        M = len(dh)
        T, N, G = gx.shape
        D, R = sum(dh), sum(rd)
        if G != 4 * D:
            raise RuntimeError(f"gx last dim {G} != 4*sum(dh) = {4 * D}")
        gx = gx.contiguous()
        weights = tuple(w.detach().contiguous() for w in weights)
        U, V = weights[0:M], weights[M:2 * M]
        Watt, batt = weights[2 * M], weights[2 * M + 1]
        Wr, br = weights[2 * M + 2:3 * M + 2], weights[3 * M + 2:4 * M + 2]
        Wf1, bf1, Wf2, bf2 = weights[4 * M + 2:4 * M + 6]
        desc = _lib.make_desc(T, N, dh, rd, map_h, 4, rows_per_cta)
        if prepacked is not None:           # mab_prepack: the images were built on a side stream while the encoders ran
            packed, ready = prepacked
            torch.cuda.current_stream(gx.device).wait_event(ready)
        else:
            packed = _pack(desc, weights, M, gx.device)
        work = _workspace(desc, gx.device)
        new = lambda *s: torch.empty(*s, device=gx.device, dtype=torch.float32)
        hz, u = new(T, N, 2 * D), new(T, N, map_h)
        need_grad = any(ctx.needs_input_grad)
        if drop_mask is not None:
            drop_mask = drop_mask.contiguous()
        if need_grad:
            sC = new(T, N, D)
            st = _lib.mab_alloc_stash(desc, gx.device)      # private piece-major stash of the kernel pair
            sCp, sG, sE, sMS, sP = st["sCp"], st["sG"], st["sE"], st["sMS"], st["sP"]
        else:
            sC = sCp = sG = sE = sMS = sP = None
        _timed("fwd", _lib.mab_fwd, desc, packed, gx, drop_mask, hz, u, sC, sCp, sG, sE, sMS, sP, work)
        launch_counter["fwd"] += 1
        # z_t = fc.3(u_t) for all steps at once, written into the z half of hz (HybridRNN_ATV.py:129)
        linear_into(u.view(T * N, map_h), Wf2, bf2, hz.view(T * N, 2 * D)[:, D:])
        if need_grad:
            ctx.save_for_backward(packed, hz, u, sC, sCp, sG, sE, sMS, sP, *weights)
            ctx.drop_mask = drop_mask
            ctx.dims = dims[:4]
        return hz

    @staticmethod
    def backward(ctx, dhz: torch.Tensor):
        packed, hz, u, sC, sCp, sG, sE, sMS, sP, *weights = ctx.saved_tensors
        dh, rd, map_h, rows_per_cta = ctx.dims
        M = len(dh)
        T, N, _ = hz.shape
        D, R, G = sum(dh), sum(rd), 4 * sum(dh)
        TN = T * N
        U, V = weights[0:M], weights[M:2 * M]
        Watt, batt = weights[2 * M], weights[2 * M + 1]
        Wr, br = weights[2 * M + 2:3 * M + 2], weights[3 * M + 2:4 * M + 2]
        Wf1, bf1, Wf2, bf2 = weights[4 * M + 2:4 * M + 6]
        desc = _lib.make_desc(T, N, dh, rd, map_h, 4, rows_per_cta)
        new = lambda *s: torch.empty(*s, device=hz.device, dtype=torch.float32)
        dhz = dhz.contiguous()
        dz_head = dhz.view(TN, 2 * D)[:, D:]
        duz = mm_nn(dz_head, Wf2)                              # the head's dL/dz pulled through fc.3: [TN, map_h]
        dgx, de, dup = new(T, N, G), new(T, N, G), new(T, N, map_h)
        att = new(T, N, G)     # attended = a * cs (HybridRNN_ATV.py:125), regrouped per modality head-major (lines 126-128) by the kernel
        _timed("bwd", _lib.mab_bwd, desc, packed, dhz, duz.view(T, N, map_h), ctx.drop_mask, sCp, sG, sE, sMS, sP, u,
               dgx, de, dup, att, _workspace(desc, hz.device))
        launch_counter["bwd"] += 1

        # ---- time-parallel weight-gradient products (fp32-accurate; allow_tf32 stays off) ----
        # z never has to be touched here: with z_{t-1} = u_{t-1} Wf2^T + bf2 and X = ds[1:]^T [h_{t-1} | u_{t-1}]  (K = (T-1) N)
        #   dU_m = block of X_h,     dV = X_u Wf2^T + colsum(ds[1:]) bf2^T,
        #   dWf2 = dz_head^T u + Vcat^T X_u,     dbf2 = colsum(dz_head) + Vcat^T colsum(ds[1:])
        # (the carried part of dL/dz_t is ds_{t+1} Vcat, and its product with u_t is Vcat^T X_u again).
        u2 = u.view(TN, map_h)
        Vcat = torch.cat(list(V), dim=0)                        # [G, D], native gate order = column order of dgx
        if T > 1:
            ds1 = dgx[1:].reshape(-1, G)
            Xh = mm_tn(ds1, hz[:-1].reshape(-1, 2 * D)[:, :D])  # [G, D]
            Xu = mm_tn(ds1, u2[:TN - N])                        # [G, map_h]
            cs1 = colsum(ds1)                                   # [G]
        else:
            Xh, Xu, cs1 = dgx.new_zeros(G, D), dgx.new_zeros(G, map_h), dgx.new_zeros(G)
        gVfull = mm_nt(Xu, Wf2) + torch.outer(cs1, bf2)         # [G, D]
        gU: List[torch.Tensor] = []
        gV: List[torch.Tensor] = []
        o = 0
        for m in range(M):
            gU.append(Xh[4 * o:4 * o + 4 * dh[m], o:o + dh[m]])
            gV.append(gVfull[4 * o:4 * o + 4 * dh[m]])
            o += dh[m]
        gWf2 = mm_tn(dz_head, u2) + mm_tn(Vcat, Xu)             # [D, map_h]
        gbf2 = colsum(dz_head) + (Vcat * cs1[:, None]).sum(0)
        de2, c2 = de.view(TN, G), sC.view(TN, D)
        gWatt, gbatt = mm_tn(de2, c2), colsum(de2)
        # reduce_m / fc.0: with r = attended Wr^T + br and dr = dup Wf1, one product Y = dup^T attended [map_h, G] gives both:
        #   dWr_m = Wf1_m^T Y_m,  dbr_m = Wf1_m^T colsum(dup),  dWf1[:, m] = Y_m Wr_m^T + colsum(dup) br_m^T,  dbf1 = colsum(dup)
        att2, dup2 = att.view(TN, G), dup.view(TN, map_h)
        Y, cdup = mm_tn(dup2, att2), colsum(dup2)
        gWr, gbr, gWf1_parts = [], [], []
        o = ro = 0
        for m in range(M):
            Ym, Wf1m = Y[:, 4 * o:4 * o + 4 * dh[m]], Wf1[:, ro:ro + rd[m]]
            gWr.append(mm_tn(Wf1m, Ym))
            gbr.append((Wf1m * cdup[:, None]).sum(0))
            gWf1_parts.append(mm_nt(Ym, Wr[m]) + torch.outer(cdup, br[m]))
            o += dh[m]
            ro += rd[m]
        gWf1, gbf1 = torch.cat(gWf1_parts, dim=1), cdup
        grads = (*gU, *gV, gWatt, gbatt, *gWr, *gbr, gWf1, gbf1, gWf2, gbf2)
        return (dgx, None, None, *grads)


def mab_recurrence(gx: torch.Tensor, drop_mask: Optional[torch.Tensor], dh: Sequence[int], rd: Sequence[int],
                   map_h: int, weights: Sequence[torch.Tensor], rows_per_cta: int = 0, prepacked=None) -> torch.Tensor:
    if not gx.is_cuda:
        raise RuntimeError("lsthm_b200: the recurrence runs on a CUDA device only (no CPU fallback)")
    return MabRecurrenceFn.apply(gx, drop_mask, (tuple(dh), tuple(rd), int(map_h), int(rows_per_cta), prepacked), *weights)
