"""torch.autograd glue for the fused encoder self-attention kernels (lsthm_attn_fwd/bwd): takes the fused
projection output qkv[B, L, 3*H*40] and returns the concatenated heads [B, L, H*40]
(model/encoder.py:38-53: split heads, ScaledDotProductAttention, merge heads)."""
from __future__ import annotations

import torch

from . import _lib

launches = {"attn": 0}
# bench.py sets this to a list to collect (start_event, end_event, useful FLOPs) per attention launch
events = None


def _timed(flops, fn, *args):
    launches["attn"] += 1
    if events is None:
        return fn(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(*args)
    e1.record()
    events.append((e0, e1, flops))


class FusedSelfAttentionFn(torch.autograd.Function):
    """qkv: [B, L, 3*H*40] (batch-major) or, with ``time_major``, [L, B, 3*H*40] — the reference's own activation
    layout, read in place through the kernels' row strides (no permute copy).  The last dim may be row-strided."""

    @staticmethod
    def forward(ctx, qkv: torch.Tensor, n_head: int, scale: float, p_drop: float, seed: int, time_major: bool):
        W = qkv.shape[-1]
        B, L = (qkv.shape[1], qkv.shape[0]) if time_major else (qkv.shape[0], qkv.shape[1])
        HD = W // 3
        qkv = qkv.contiguous()
        out = torch.empty(*qkv.shape[:2], HD, device=qkv.device, dtype=torch.float32)
        lse = torch.empty(B * n_head, L, device=qkv.device, dtype=torch.float32)
        d = _lib.make_attn_desc(B, L, n_head, W, W, W, HD, scale, p_drop, seed, time_major=time_major)
        q, k, v = qkv[:, :, :HD], qkv[:, :, HD:2 * HD], qkv[:, :, 2 * HD:]
        _timed(4.0 * B * n_head * L * L * 40, _lib.attn_fwd, d, q, k, v, out, lse)     # QK^T and PV
        ctx.save_for_backward(qkv, out, lse)
        ctx.cfg = (n_head, scale, p_drop, seed, time_major, B, L)
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        qkv, out, lse = ctx.saved_tensors
        n_head, scale, p_drop, seed, time_major, B, L = ctx.cfg
        W = qkv.shape[-1]
        HD = W // 3
        dqkv = torch.empty_like(qkv)
        d = _lib.make_attn_desc(B, L, n_head, W, W, W, HD, scale, p_drop, seed, time_major=time_major)
        q, k, v = qkv[:, :, :HD], qkv[:, :, HD:2 * HD], qkv[:, :, 2 * HD:]
        _timed(10.0 * B * n_head * L * L * 40, _lib.attn_bwd, d, q, k, v, out, lse, dout.contiguous(), dqkv[:, :, :HD],
               dqkv[:, :, HD:2 * HD], dqkv[:, :, 2 * HD:])                               # S, dPd, dV, dQ, dK
        return dqkv, None, None, None, None, None


def fused_self_attention(qkv: torch.Tensor, n_head: int, scale: float, p_drop: float = 0.0, seed: int = 0,
                         time_major: bool = False) -> torch.Tensor:
    return FusedSelfAttentionFn.apply(qkv, n_head, float(scale), float(p_drop), int(seed), bool(time_major))
