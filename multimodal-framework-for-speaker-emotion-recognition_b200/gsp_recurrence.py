"""torch.autograd glue around the GRU speaker-state cell kernels (C ABI: lsthm_gsp_* in
include/lsthm_b200.h).  Replaces ``MARN_cell.forward`` of model/lsthm_onlysp.py:156-188 (listener = 0) and of
model/lsthm_nsps.py:160-198 (listener = 1).

The input-side products (``W x`` of both LSTHM1 cells and ``weight_ih U`` of the GRU) are time-parallel and arrive as
``gx`` / ``gxs``; the kernel owns everything that is sequential.  Backward: the BPTT kernel returns the adjoints of all
gate pre-activations; the recurrent weight gradients are time-parallel products formed here.
"""
from __future__ import annotations

import torch

from . import _lib
from . import recurrence as _rec
from .mm3 import colsum, mm_tn
from .recurrence import launch_counter


class GspCellFn(torch.autograd.Function):
    """out[T,N,512] = cell(gx[T,N,2,512], gxs[T,N,384], qmask; weights).  Weight order:
    U_l,U_a, V_l,V_a, S_l,S_a, gru_s.weight_hh, gru_s.bias_hh, Wq, Wk."""

    @staticmethod
    def forward(ctx, gx, gxs, qmask, masks, opts, *weights):
        T, N = gx.shape[0], gx.shape[1]
        listener, rows_per_cta, att_p, att_seed = opts
        gx, gxs = gx.contiguous(), gxs.contiguous()
        qmask = qmask.contiguous().float()
        weights = tuple(w.detach().contiguous() for w in weights)
        U, V, S = weights[0:2], weights[2:4], weights[4:6]
        Whh, bhh, Wq, Wk = weights[6:10]
        ms, ml, ma, att_mask = (None if m is None else m.contiguous() for m in masks)
        desc = _lib.make_gsp_desc(T, N, listener, rows_per_cta, 0.0 if att_mask is not None else att_p, att_seed)
        w = _lib.make_gsp_weights(U, V, S, Whh, bhh, Wq, Wk)
        mk = _lib.make_gsp_masks(ms, ml, ma, att_mask)
        new = lambda *s: torch.empty(*s, device=gx.device, dtype=torch.float32)
        packed = new(_lib.gsp_packed_floats())
        _lib.gsp_pack(w, packed)
        launch_counter["pack"] += 1
        out = new(T, N, 512)
        need_grad = any(ctx.needs_input_grad)
        if need_grad:
            sGS, sQS, sGL, sCL = new(T, N, 4, 128), new(T, N, 128), new(T, N, 2, 512), new(T, N, 2, 128)
        else:
            sGS = sQS = sGL = sCL = None
        _rec._timed("fwd", _lib.gsp_fwd, desc, w, packed, gx, gxs, qmask, mk, out, sGS, sQS, sGL, sCL)
        launch_counter["fwd"] += 1
        if need_grad:
            ctx.save_for_backward(qmask, out, sGS, sQS, sGL, sCL, *weights)
            ctx.masks = (ms, ml, ma, att_mask)
            ctx.opts = opts
        return out

    @staticmethod
    def backward(ctx, dout):
        qmask, out, sGS, sQS, sGL, sCL, *weights = ctx.saved_tensors
        listener, rows_per_cta, att_p, att_seed = ctx.opts
        ms, ml, ma, att_mask = ctx.masks
        T, N = out.shape[0], out.shape[1]
        U, V, S = weights[0:2], weights[2:4], weights[4:6]
        Whh, bhh, Wq, Wk = weights[6:10]
        desc = _lib.make_gsp_desc(T, N, listener, rows_per_cta, 0.0 if att_mask is not None else att_p, att_seed)
        w = _lib.make_gsp_weights(U, V, S, Whh, bhh, Wq, Wk)
        mk = _lib.make_gsp_masks(ms, ml, ma, att_mask)
        new = lambda *s: torch.empty(*s, device=out.device, dtype=torch.float32)
        dGL, dGi, dGh = new(T, N, 2, 512), new(T, N, 384), new(T, N, 384)
        dWqk = new(_lib.gsp_launch_info(desc)["grid"], 2, 128)
        _rec._timed("bwd", _lib.gsp_bwd, desc, w, qmask, mk, dout.contiguous(), sGS, sQS, sGL, sCL, dGL, dGi, dGh, dWqk)
        launch_counter["bwd"] += 1
        # ---- time-parallel weight-gradient products (fp32) ----
        TN = T * N
        z_prev, hs_t = out[:-1, :, 256:384].reshape(-1, 128), out[:, :, 384:512].reshape(TN, 128)
        gU, gV, gS = [], [], []
        for c in range(2):
            ds = dGL[:, :, c]                                     # [T,N,512]
            ds1 = ds[1:].reshape(-1, 512)
            gU.append(mm_tn(ds1, out[:-1, :, c * 128:(c + 1) * 128].reshape(-1, 128)))   # h_{t-1} after dropout
            gV.append(mm_tn(ds1, z_prev))
            gS.append(mm_tn(ds.reshape(TN, 512), hs_t))
        dgh = dGh.view(TN, 384)
        gWhh = mm_tn(dgh, sQS.view(TN, 128))
        gbhh = colsum(dgh)
        g = dWqk.sum(0)
        grads = (*gU, *gV, *gS, gWhh, gbhh, g[0].view_as(Wq), g[1].view_as(Wk))
        return (dGL, dGi, None, None, None, *grads)


def gsp_cell(gx, gxs, qmask, masks, weights, listener: int, rows_per_cta: int = 0, att_p: float = 0.0, att_seed: int = 0):
    if not gx.is_cuda:
        raise RuntimeError("lsthm_b200: the GRU speaker-state cell runs on a CUDA device only (no CPU fallback)")
    return GspCellFn.apply(gx, gxs, qmask, tuple(masks), (int(listener), int(rows_per_cta), float(att_p), int(att_seed)), *weights)
