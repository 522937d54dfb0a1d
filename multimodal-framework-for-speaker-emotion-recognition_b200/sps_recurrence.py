"""torch.autograd glue around the speaker-state cell kernels (C ABI: lsthm_sps_* in
include/lsthm_b200.h).  Replaces ``MARN_cell.forward`` (model/lsthm_sps.py:156-221).

The packing permutation of ``_select_parties`` (lsthm_sps.py:238-259) depends only on ``qmask``; it is
computed here once per call with a few tensor ops (no per-row Python loop, no host sync).
Backward: the BPTT kernel returns the adjoints of all gate pre-activations; weight gradients are
time-parallel products formed here.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from . import recurrence as _rec
from .mm3 import mm_tn
from .recurrence import launch_counter


def party_plan(qmask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """qmask [T,N,2] -> (pi [T,N] int32: dialogue at packed row r, pr [T,N] int32: packed row of dialogue
    d, n0 [T] int32).  Packed order = speaker-0 dialogues (ascending id) then speaker-1 dialogues; an
    all-zero (padded) row counts as speaker 0 because argmax of [0,0] is 0 (lsthm_sps.py:177)."""
    spk1 = torch.argmax(qmask, dim=-1) == 1                       # [T,N]
    is0 = ~spk1
    n0 = is0.sum(1)
    pr = torch.where(is0, is0.cumsum(1) - 1, n0[:, None] + spk1.cumsum(1) - 1)
    pi = torch.argsort(pr, dim=1)
    return pi.to(torch.int32).contiguous(), pr.to(torch.int32).contiguous(), n0.to(torch.int32).contiguous()


class SpsCellFn(torch.autograd.Function):
    """out[T,N,512] = cell(gx[T,N,2,512], qmask; weights).  Weight order:
    U_l,U_a, V_l,V_a, S_l,S_a, Wih_q0,Wih_q1, Whh_q0,Whh_q1, bq0,bq1 (bias_ih+bias_hh), Wq, Wk."""

    @staticmethod
    def forward(ctx, gx, qmask, masks, opts, *weights):
        T, N = gx.shape[0], gx.shape[1]
        rows_per_cta, att_p, att_seed = opts
        gx = gx.contiguous()
        qmask = qmask.contiguous().float()
        weights = tuple(w.detach().contiguous() for w in weights)
        U, V, S, Wih, Whh, bq = (weights[0:2], weights[2:4], weights[4:6], weights[6:8], weights[8:10], weights[10:12])
        Wq, Wk = weights[12], weights[13]
        pi, pr, n0 = party_plan(qmask)
        mq0, mq1, ml, ma, att_mask = (None if m is None else m.contiguous() for m in masks)
        desc = _lib.make_sps_desc(T, N, rows_per_cta, 0.0 if att_mask is not None else att_p, att_seed)
        w = _lib.make_sps_weights(U, V, S, Wih, Whh, bq, Wq, Wk)
        mk = _lib.make_sps_masks(mq0, mq1, ml, ma, att_mask)
        new = lambda *s: torch.empty(*s, device=gx.device, dtype=torch.float32)
        packed = new(_lib.sps_packed_floats())
        _lib.sps_pack(w, packed)
        launch_counter["pack"] += 1
        ws = new(_lib.sps_workspace_floats(desc))
        out = new(T, N, 512)
        need_grad = any(ctx.needs_input_grad)
        if need_grad:
            sGQ, sGL = new(T, N, 2, 512), new(T, N, 2, 512)
            sCQ, sHQ, sXQ, sCL, sHL = (new(T, N, 2, 128) for _ in range(5))
        else:
            sGQ = sGL = sCQ = sHQ = sXQ = sCL = sHL = None
        _rec._timed("fwd", _lib.sps_fwd, desc, w, packed, gx, qmask, pi, n0, mk, ws, out, sGQ, sCQ, sHQ, sXQ, sGL, sCL, sHL)
        launch_counter["fwd"] += 1
        if need_grad:
            ctx.save_for_backward(qmask, pi, pr, n0, out, sGQ, sCQ, sHQ, sXQ, sGL, sCL, sHL, ws, *weights)
            ctx.masks = (mq0, mq1, ml, ma, att_mask)
            ctx.opts = opts
        return out

    @staticmethod
    def backward(ctx, dout):
        qmask, pi, pr, n0, out, sGQ, sCQ, sHQ, sXQ, sGL, sCL, sHL, ws, *weights = ctx.saved_tensors
        rows_per_cta, att_p, att_seed = ctx.opts
        mq0, mq1, ml, ma, att_mask = ctx.masks
        T, N = out.shape[0], out.shape[1]
        U, V, S, Wih, Whh, bq = (weights[0:2], weights[2:4], weights[4:6], weights[6:8], weights[8:10], weights[10:12])
        Wq, Wk = weights[12], weights[13]
        desc = _lib.make_sps_desc(T, N, rows_per_cta, 0.0 if att_mask is not None else att_p, att_seed)
        w = _lib.make_sps_weights(U, V, S, Wih, Whh, bq, Wq, Wk)
        mk = _lib.make_sps_masks(mq0, mq1, ml, ma, att_mask)
        new = lambda *s: torch.empty(*s, device=out.device, dtype=torch.float32)
        dGL, dGQ = new(T, N, 2, 512), new(T, N, 2, 512)
        grid = _lib.sps_launch_info(desc)["grid"]
        dWqk = new(grid, 2, 128)
        _rec._timed("bwd", _lib.sps_bwd, desc, w, qmask, pi, pr, n0, mk, dout.contiguous(), sGQ, sCQ, sGL, sCL, ws, dGL, dGQ,
                    dWqk)
        launch_counter["bwd"] += 1
        # ---- time-parallel weight-gradient products (fp32) ----
        TN = T * N
        z_prev, hq_t = out[:-1, :, 256:384].reshape(-1, 128), out[:, :, 384:512].reshape(TN, 128)
        gU, gV, gS, gWih, gWhh, gbq = [], [], [], [], [], []
        for c in range(2):
            ds = dGL[:, :, c]                                     # [T,N,512]
            ds1 = ds[1:].reshape(-1, 512)
            gU.append(mm_tn(ds1, sHL[:-1, :, c].reshape(-1, 128)))   # h_{t-1} after dropout (lsthm_sps.py:211,213)
            gV.append(mm_tn(ds1, z_prev))
            gS.append(mm_tn(ds.reshape(TN, 512), hq_t))
            dg = dGQ[:, :, c]
            gWih.append(mm_tn(dg.reshape(TN, 512), sXQ[:, :, c].reshape(TN, 128)))
            gWhh.append(mm_tn(dg[1:].reshape(-1, 512), sHQ[:-1, :, c].reshape(-1, 128)))
            gbq.append(dg.reshape(TN, 512).sum(0))
        g = dWqk.sum(0)
        grads = (*gU, *gV, *gS, *gWih, *gWhh, *gbq, g[0].view_as(Wq), g[1].view_as(Wk))
        return (dGL, None, None, None, *grads)


def sps_cell(gx, qmask, masks, weights, rows_per_cta: int = 0, att_p: float = 0.0, att_seed: int = 0):
    if not gx.is_cuda:
        raise RuntimeError("lsthm_b200: the speaker-state cell runs on a CUDA device only (no CPU fallback)")
    return SpsCellFn.apply(gx, qmask, tuple(masks), (int(rows_per_cta), float(att_p), int(att_seed)), *weights)
