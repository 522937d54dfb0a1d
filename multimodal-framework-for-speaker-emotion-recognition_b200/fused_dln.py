"""torch.autograd glue for the fused  LayerNorm(dropout(y + bias) + residual)  kernels (lsthm_dln_fwd/bwd), the tail
of both halves of the reference's EncoderLayer (model/encoder.py:54-58 and :106-112)."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

launches = {"dln": 0}


class DropResLayerNormFn(torch.autograd.Function):
    """y, res: 2-D row matrices [R, d] (unit inner stride, row strides multiples of 4; res may be row-strided)."""

    @staticmethod
    def forward(ctx, y, bias, res, gamma, beta, eps: float, p_drop: float, seed: int):
        R, d = y.shape
        need = any(ctx.needs_input_grad)
        out = torch.empty(R, d, device=y.device, dtype=torch.float32)
        v = torch.empty(R, d, device=y.device, dtype=torch.float32) if need else None
        desc = _lib.make_dln_desc(R, d, eps, p_drop, seed)
        _lib.dln_fwd(desc, y, None if bias is None else bias.contiguous(), res, gamma.contiguous(), beta.contiguous(), v, out)
        launches["dln"] += 1
        if need:
            ctx.save_for_backward(v, gamma)
            ctx.cfg = (eps, p_drop, seed, bias is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        v, gamma = ctx.saved_tensors
        eps, p_drop, seed, has_bias = ctx.cfg
        R, d = v.shape
        if dout.stride(1) != 1 or dout.stride(0) % 4 or dout.data_ptr() % 16:
            dout = dout.contiguous()
        new = lambda *s: torch.empty(*s, device=v.device, dtype=torch.float32)
        dres = new(R, d)
        dy = new(R, d) if p_drop > 0 else None
        dgamma, dbeta = new(d), new(d)
        dbias = new(d) if has_bias else None
        _lib.dln_bwd(_lib.make_dln_desc(R, d, eps, p_drop, seed), dout, v, gamma.contiguous(), dy, dres, dgamma, dbeta, dbias)
        launches["dln"] += 2
        return (dres if dy is None else dy), dbias, dres, dgamma, dbeta, None, None, None


def drop_res_layer_norm(y: torch.Tensor, bias: Optional[torch.Tensor], res: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                        eps: float, p_drop: float = 0.0, seed: int = 0) -> torch.Tensor:
    return DropResLayerNormFn.apply(y, bias, res, gamma, beta, float(eps), float(p_drop), int(seed))
