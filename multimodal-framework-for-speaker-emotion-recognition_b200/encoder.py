"""Host-side mirror of the reference's per-modality utterance encoder (model/encoder.py) — the first "next"
row of SURVEY.md §8(f).

Parameter names, shapes and construction order equal the reference's (encoder.py:10-25, 92-99, 124-127) so
that state_dicts load both ways and a fixed seed yields identical initial weights.  On a CUDA device the whole
layer runs on our own kernels over a 2-D row view of the activations (``EncoderLayer._fast``): fused QKV
projection and the other Linear layers on ``lsthm_gemm3`` (tcgen05), self-attention on ``lsthm_attn_*`` (tcgen05,
reads time-major rows in place), and both ``LayerNorm(dropout(.) + residual)`` tails on ``lsthm_dln_*``.  The
module-by-module path below it is the same arithmetic in PyTorch ops and is what runs when a test swaps the
dropout modules for a mask tape, or on shapes the kernels do not cover.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .fused_attention import fused_self_attention
from .fused_dln import drop_res_layer_norm
from .mm3 import linear3, rows_view, _LinearTC
from . import mm3

# There is ONE CUDA path (EncoderLayer._fast: own tcgen05 GEMM + fused attention + fused dropout/residual/LayerNorm).  The
# module-by-module torch expressions below are the reference semantics for what no kernel covers: CPU tensors and fp64
# (the tests' truth runs), dropout modules swapped for a mask tape (train-mode parity), an explicit mask,
# or a caller that wants the [B,H,L,L] attention weights (``need_weights``).


class ScaledDotProductAttention(nn.Module):
    """softmax(q k^T / temperature) v with dropout on the weights; never masked in this repo
    (encoder.py:71-86, call site encoder.py:131 passes mask=None)."""

    def __init__(self, temperature: float, attn_dropout: float = 0.1):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(attn_dropout)

    def forward(self, q, k, v, mask=None):
        scores = torch.matmul(q / self.temperature, k.transpose(-2, -1))
        if mask is not None:
            scores = scores.masked_fill(mask == 0, -1e9)
        attn = self.dropout(torch.softmax(scores, dim=-1))
        return torch.matmul(attn, v), attn


class MultiHeadAttention(nn.Module):
    def __init__(self, n_head, d_model, d_model2, d_k, d_v, dropout=0.1):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(d_model2, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(d_model2, n_head * d_v, bias=False)
        self.fc = nn.Linear(n_head * d_v, d_model, bias=False)
        self.attention = ScaledDotProductAttention(temperature=d_k ** 0.5)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.need_weights = False      # True: return the [B,H,L,L] attention weights as the reference does (explicit softmax path)

    def forward(self, q, k, v, mask=None):
        B, Lq, Lk = q.size(0), q.size(1), k.size(1)
        H, dk, dv = self.n_head, self.d_k, self.d_v
        res = q
        if (not self.need_weights and q is k and k is v and mask is None and q.is_cuda and q.dtype == torch.float32 and Lq <= 128
                and dk == 40 and dv == 40 and type(self.attention.dropout) is nn.Dropout and q.shape[-1] % 4 == 0):
            w_qkv = torch.cat([self.w_qs.weight, self.w_ks.weight, self.w_vs.weight], dim=0)
            qkv = linear3(q, w_qkv)                                            # one projection GEMM instead of three
            p = self.attention.dropout.p if self.training else 0.0
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
            ctx = fused_self_attention(qkv, H, 1.0 / self.attention.temperature, p, seed)
            out = self.layer_norm(self.dropout(linear3(ctx, self.fc.weight)) + res)
            return out, None
        qh = linear3(q, self.w_qs.weight).view(B, Lq, H, dk).transpose(1, 2)
        kh = linear3(k, self.w_ks.weight).view(B, Lk, H, dk).transpose(1, 2)
        vh = linear3(v, self.w_vs.weight).view(B, Lk, H, dv).transpose(1, 2)
        ctx, attn = self.attention(qh, kh, vh, mask=None if mask is None else mask.unsqueeze(1))
        ctx = ctx.transpose(1, 2).reshape(B, Lq, H * dv)
        out = self.layer_norm(self.dropout(linear3(ctx, self.fc.weight)) + res)
        return out, attn


class PositionwiseFeedForward(nn.Module):
    def __init__(self, d_in, d_hid, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Linear(d_in, d_hid)
        self.w_2 = nn.Linear(d_hid, d_in)
        self.layer_norm = nn.LayerNorm(d_in, eps=1e-6)
        self.dropout = nn.Dropout(dropout)
        self.fc = nn.Linear(d_in, 100)  # registered but never applied (encoder.py:99,111): grad stays None

    def forward(self, x):
        w1, b1, w2 = self.w_1.weight, self.w_1.bias, self.w_2.weight
        pad = (-w1.shape[0]) % 4
        if pad and x.is_cuda:
            # d_inner = 50 (HybridRNN_AT/ATV) is not a multiple of 4: zero-pad the hidden width so both products
            # run on the tensor-core GEMM (padded units are relu(0) = 0 and meet zero columns of w_2)
            w1, b1, w2 = F.pad(w1, (0, 0, 0, pad)), F.pad(b1, (0, pad)), F.pad(w2, (0, pad))
        h = F.relu(linear3(x, w1, b1))
        return self.layer_norm(self.dropout(linear3(h, w2, self.w_2.bias)) + x)


def _seed(p: float) -> int:
    return int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0


class EncoderLayer(nn.Module):
    def __init__(self, d_model, d_inner, n_head, d_k, d_v, dropout=0.1):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_model, d_k, d_v, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, dropout=dropout)

    def _fast_ok(self, x: torch.Tensor, mask) -> bool:
        a, f = self.slf_attn, self.pos_ffn
        plain = all(type(m) is nn.Dropout for m in (a.dropout, a.attention.dropout, f.dropout))
        return (not a.need_weights and mask is None and x.is_cuda and x.dtype == torch.float32 and x.dim() == 3
                and x.shape[1] <= 128 and a.d_k == 40 and a.d_v == 40 and x.shape[-1] % 4 == 0 and x.shape[-1] <= 512
                and plain and x.data_ptr() % 16 == 0)

    def _fast(self, x: torch.Tensor) -> torch.Tensor:
        """The whole layer on 2-D rows; ``x`` is [B, L, d] logically, in whatever row order its storage has."""
        a, f = self.slf_attn, self.pos_ffn
        B, L, d = x.shape
        rv = rows_view(x)
        if rv is None:
            x = x.contiguous()
            rv = rows_view(x)
        rows, time_major = rv                                   # rows [R, d]; time_major: row = i*B + b
        train = self.training
        H, W = a.n_head, 3 * a.n_head * 40
        w_qkv = torch.cat([a.w_qs.weight, a.w_ks.weight, a.w_vs.weight], dim=0)
        qkv = _LinearTC.apply(rows, w_qkv, None, False)         # one projection GEMM instead of three (encoder.py:38-40)
        p_att = a.attention.dropout.p if train else 0.0
        qkv3 = qkv.view(L, B, W) if time_major else qkv.view(B, L, W)
        ctx = fused_self_attention(qkv3, H, 1.0 / a.attention.temperature, p_att, _seed(p_att), time_major)
        y = _LinearTC.apply(ctx.view(B * L, H * 40), a.fc.weight, None, False)
        p1 = a.dropout.p if train else 0.0
        x1 = drop_res_layer_norm(y, None, rows, a.layer_norm.weight, a.layer_norm.bias, a.layer_norm.eps, p1, _seed(p1))
        w1, b1, w2 = f.w_1.weight, f.w_1.bias, f.w_2.weight
        pad = (-w1.shape[0]) % 4
        if pad:   # d_inner = 50 (HybridRNN_AT/ATV): zero-pad the hidden width to a multiple of 4 (relu(0) = 0 meets zero columns)
            w1, b1, w2 = F.pad(w1, (0, 0, 0, pad)), F.pad(b1, (0, pad)), F.pad(w2, (0, pad))
        h = _LinearTC.apply(x1, w1, b1, True)                   # relu(w_1 x + b) with the ReLU in the GEMM epilogue
        y2 = _LinearTC.apply(h, w2, None, False)                # w_2's bias is added by the fused tail (its gradient comes from there)
        p2 = f.dropout.p if train else 0.0
        out = drop_res_layer_norm(y2, f.w_2.bias, x1, f.layer_norm.weight, f.layer_norm.bias, f.layer_norm.eps, p2, _seed(p2))
        return out.view(L, B, d).transpose(0, 1) if time_major else out.view(B, L, d)

    def forward(self, enc_input, slf_attn_mask=None):
        if self._fast_ok(enc_input, slf_attn_mask):
            return self._fast(enc_input), None
        y, attn = self.slf_attn(enc_input, enc_input, enc_input, mask=slf_attn_mask)
        return self.pos_ffn(y), attn
