"""Drop-in for the reference's ``model/lsthm_sps.py``: ``MARN1_sps(n_classes)`` ->
``forward(x[L,B,1124], qmask[L,B,2], umask[B,L]) -> (log-probs [B*L, C], x_l, x_a)``.

Same constructor, parameter names / shapes / registration order (including the never-used ones:
``crossatt_*.Wv`` of the in-cell attention, all of ``marn_cell_*.crossatt_a2l`` and ``lstm_s``;
SURVEY.md F8) and default-init RNG order as the reference (lsthm_sps.py:11-26, 47-57, 75-86, 103-114,
132-154, 298-346).  ``MARN_cell.forward`` (156-221) runs as one fused CUDA kernel per direction;
the encoders and the sequence-level cross attention (88-129) run on the fused attention / GEMM kernels.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .encoder import EncoderLayer
from .streams import fork_join, state_without_streams
from .mm3 import linear3, linear_cat
from .seq_attention import fused_ok, seq_cross_attention
from .sps_recurrence import sps_cell


class LSTHM1(nn.Module):
    """Parameter container of one LSTHM1 cell (lsthm_sps.py:11-19); arithmetic is in the kernel."""

    def __init__(self, cell_size, in_size, hybrid_in_size, speaker_dim):
        super().__init__()
        self.cell_size, self.in_size = cell_size, in_size
        self.W = nn.Linear(in_size, 4 * cell_size)
        self.U = nn.Linear(cell_size, 4 * cell_size)
        self.V = nn.Linear(hybrid_in_size, 4 * cell_size)
        self.S = nn.Linear(speaker_dim, 4 * cell_size)

    def gate_input(self, x):
        """W x plus the four biases, for all steps at once (time-parallel part of lines 29-34)."""
        return linear3(x, self.W.weight, self.W.bias + self.U.bias + self.V.bias + self.S.bias)

    def forward(self, x, ctm, htm, ztm, speaker_affine):
        """One step in the reference's signature (lsthm_sps.py:28-44), for callers that drive a cell by hand; plain tensor
        expressions — the model never calls it (its cells run inside the fused kernels)."""
        s = self.W(x) + self.U(htm) + self.V(ztm) + self.S(speaker_affine)
        d = self.cell_size
        f, i, o = torch.sigmoid(s[:, :d]), torch.sigmoid(s[:, d:2 * d]), torch.sigmoid(s[:, 2 * d:3 * d])
        c = f * ctm + i * torch.tanh(s[:, 3 * d:])
        return c, torch.tanh(c) * o


class CrossAttention(nn.Module):
    """In-cell rank-1 attention parameters (lsthm_sps.py:47-57); evaluated inside the kernel."""

    def __init__(self, attn_dropout=0.2):
        super().__init__()
        self.dh = 128
        self.Wq = nn.Parameter(torch.ones(self.dh).unsqueeze(0))
        self.Wk = nn.Parameter(torch.ones(self.dh).unsqueeze(0))
        self.Wv = nn.Parameter(torch.ones(self.dh).unsqueeze(0))
        self.dropout = nn.Dropout(attn_dropout)

    def forward(self, x_1, x_2):
        """One call in the reference's signature (lsthm_sps.py:59-72), in the collapsed rank-1 form the kernels use
        (a_i = x1_i (Wq.x2)/sqrt(128); out_i = sum_j dropout(softmax_j(a_i Wk_j)) x2_j); never called by the model."""
        a = x_1 * ((x_2 * self.Wq).sum(-1, keepdim=True) / self.dh ** 0.5)              # [N, D]
        w = self.dropout(torch.softmax(a.unsqueeze(-1) * self.Wk.view(1, 1, -1), dim=-1))   # [N, D, D]
        return torch.matmul(w, x_2.unsqueeze(-1)).squeeze(-1)


class _SeqCrossAttention(nn.Module):
    """Dense, unmasked attention over the utterances of a dialogue (lsthm_sps.py:88-101, 116-129)."""

    def __init__(self, d_q, d_kv, attn_dropout=0.2, dk=128, dv=128):
        super().__init__()
        self.dh, self.dk, self.dv = 100, dk, dv
        self.Wq = nn.Parameter(torch.ones(d_q, self.dk))
        self.Wk = nn.Parameter(torch.ones(d_kv, self.dk))
        self.Wv = nn.Parameter(torch.ones(d_kv, self.dv))
        self.dropout = nn.Dropout(attn_dropout)

    def forward(self, x_1, x_2, s_1=None, s_2=None):
        """``s_1`` / ``s_2``: optional learnable scalars multiplying x_1 / x_2 (``self.w * x_l`` etc., lsthm_sps.py:377-383).
        On the fused path they are folded into the projection weights — (s x) W = x (s W): a [d, 128] product instead of an
        [L, B, d] one, forward and backward."""
        if type(self.dropout) is nn.Dropout and fused_ok(x_1, x_2, self.dk, self.dv):
            # own kernels: six-term tensor-core projections + fused attention core (seq_attention.py)
            p = self.dropout.p if self.training else 0.0
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
            Wq = self.Wq if s_1 is None else self.Wq * s_1
            Wk, Wv = (self.Wk, self.Wv) if s_2 is None else (self.Wk * s_2, self.Wv * s_2)
            return seq_cross_attention(x_1, x_2, Wq, Wk, Wv, p, seed)
        # explicit form: CPU / fp64 truth runs of the tests, and train-mode parity runs that drive the dropout from a mask tape
        if s_1 is not None:
            x_1 = s_1 * x_1
        if s_2 is not None:
            x_2 = s_2 * x_2
        a, b = x_1.permute(1, 0, 2), x_2.permute(1, 0, 2)
        q, k, v = a @ self.Wq, b @ self.Wk, b @ self.Wv
        w = self.dropout(torch.softmax((q / self.dk ** 0.5) @ k.transpose(1, 2), dim=-1))
        return (w @ v).permute(1, 0, 2)


class CrossAttention2(_SeqCrossAttention):
    def __init__(self, dh, dk, dv, attn_dropout=0.2):
        super().__init__(100, 100, attn_dropout)


class CrossAttention3(_SeqCrossAttention):
    def __init__(self, dh, dk, dv, attn_dropout=0.2):
        super().__init__(100, 128, attn_dropout)


class MARN_cell(nn.Module):
    def __init__(self, dh_l, dh_a, d_l, d_a, dropout=0.5):
        super().__init__()
        self.crossatt_l2a = CrossAttention()
        self.crossatt_a2l = CrossAttention()          # never used by the reference either (line 216)
        self.dh_l, self.dh_a, self.dh_q, self.d_l, self.d_a, self.dh_s = dh_l, dh_a, dh_l, d_l, d_a, 128
        self.lsthm_l = LSTHM1(dh_l, d_l, dh_l, self.dh_s)
        self.lsthm_a = LSTHM1(dh_a, d_a, dh_l, self.dh_s)
        self.lstm_q0 = nn.LSTMCell(self.dh_s, self.dh_s)
        self.lstm_q1 = nn.LSTMCell(self.dh_s, self.dh_s)
        self.lstm_s = nn.LSTMCell(self.dh_s, self.dh_s)   # never used (line 153)
        self.dropout = nn.Dropout(dropout)
        self.rows_per_cta = 0
        self.mask_override = None      # test hook: (mq0, mq1, ml, ma, att_mask) dropout mask tape

    def cell_weights(self):
        l, a, q0, q1 = self.lsthm_l, self.lsthm_a, self.lstm_q0, self.lstm_q1
        return [l.U.weight, a.U.weight, l.V.weight, a.V.weight, l.S.weight, a.S.weight,
                q0.weight_ih, q1.weight_ih, q0.weight_hh, q1.weight_hh,
                q0.bias_ih + q0.bias_hh, q1.bias_ih + q1.bias_hh, self.crossatt_l2a.Wq, self.crossatt_l2a.Wk]

    def forward(self, x, x_l, x_a, qmask):
        T, N, _ = x_l.shape
        l, a = self.lsthm_l, self.lsthm_a
        gx = linear_cat([x_l, x_a], [l.W.weight, a.W.weight],
                        [l.W.bias + l.U.bias + l.V.bias + l.S.bias, a.W.bias + a.U.bias + a.V.bias + a.S.bias]).view(T, N, 2, 512)
        att_p, seed = 0.0, 0
        if self.mask_override is not None:
            masks = self.mask_override
        elif self.training:
            p = self.dropout.p
            draw = lambda: torch.empty(T, N, 128, device=gx.device).bernoulli_(1 - p).mul_(1 / (1 - p)) if p > 0 else None
            masks = (draw(), draw(), draw(), draw(), None)
            att_p = self.crossatt_l2a.dropout.p
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        else:
            masks = (None,) * 5
        return sps_cell(gx, qmask, masks, self.cell_weights(), self.rows_per_cta, att_p, seed)


class _ReverseSeq(torch.autograd.Function):
    """One kernel each way; the map is its own adjoint (``lsthm_reverse_seq``)."""

    @staticmethod
    def forward(ctx, X, lens):
        ctx.save_for_backward(lens)
        return _lib.reverse_seq(X.contiguous(), lens)

    @staticmethod
    def backward(ctx, g):
        (lens,) = ctx.saved_tensors
        return _lib.reverse_seq(g.contiguous(), lens), None


def reverse_seq(X: torch.Tensor, umask: torch.Tensor) -> torch.Tensor:
    """Per-dialogue flip over its own length with zero padding (MARN1_sps._reverse_seq, lsthm_sps.py:396-410): no Python loop
    over the batch, no host sync.  fp32 CUDA tensors of even width go through one fused kernel (forward and backward); the
    index form below is for CPU / fp64 tensors of the tests' truth runs."""
    L = X.shape[0]
    lens = umask.sum(1)
    if X.is_cuda and X.dtype == torch.float32 and X.dim() == 3 and X.shape[2] % 2 == 0:
        return _ReverseSeq.apply(X, lens.to(torch.int32))
    lens = lens.long()                                           # [B]
    src = lens[None, :] - 1 - torch.arange(L, device=X.device)[:, None]   # [L,B]
    valid = (src >= 0).to(X.dtype).unsqueeze(-1)
    idx = src.clamp(min=0).unsqueeze(-1).expand(-1, -1, X.shape[2])
    return X.gather(0, idx) * valid


class MARN1_sps(nn.Module):
    def __getstate__(self):
        return state_without_streams(self)          # cached CUDA streams are not part of the module's state

    def __init__(self, n_classes):
        super().__init__()
        self.d_l, self.d_a, self.d_r = 100, 100, 1024
        self.dh_l, self.dh_a, self.dh_sp, self.dh_li = 128, 128, 128, 128
        self.total_h_dim = self.dh_l + self.dh_a
        self.linear_in = nn.Linear(self.d_r, self.d_l)
        self.marn_cell_f = MARN_cell(self.dh_l, self.dh_a, self.d_l, self.d_a)
        self.marn_cell_b = MARN_cell(self.dh_l, self.dh_a, self.d_l, self.d_a)
        final_out = 2 * (self.total_h_dim + self.dh_l + self.dh_l) + self.dh_l + self.dh_a
        self.fc = nn.Sequential(nn.Linear(final_out, self.d_l), nn.ReLU(), nn.Dropout(0.5))
        self.nn_out = nn.Sequential(nn.Linear(self.d_l, 32), nn.ReLU(), nn.Dropout(0.5), nn.Linear(32, n_classes))
        self.dropout_rec = nn.Dropout(0.5)
        self.encoder_l = EncoderLayer(100, 40, 8, 40, 40)
        self.encoder_a = EncoderLayer(100, 40, 8, 40, 40)
        self.crossatt_l2a = CrossAttention2(self.d_l, self.dh_l, self.dh_l)
        self.crossatt_a2l = CrossAttention2(self.d_a, self.dh_a, self.dh_a)
        self.crossatt_l2a_1 = CrossAttention3(self.dh_l, self.d_l, self.d_l)
        self.crossatt_a2l_1 = CrossAttention3(self.dh_a, self.d_a, self.d_a)
        self.w = nn.Parameter(torch.ones(1))
        self.v = nn.Parameter(torch.ones(1))
        self.v1 = nn.Parameter(torch.ones(1))
        self.v2 = nn.Parameter(torch.ones(1))
        self.concurrent_encoders = True           # text / audio encoder chains on two CUDA streams (same results bit for bit)

    def forward(self, x, qmask, umask):
        x_l = linear3(x[:, :, :self.d_r].permute(1, 0, 2), self.linear_in.weight, self.linear_in.bias)
        x_a = x[:, :, self.d_r:self.d_r + self.d_a].permute(1, 0, 2)
        # enc(x + enc(x)) per modality (lsthm_sps.py:356-361); text and audio are independent up to the cells: on a CUDA device
        # the two chains run on two streams, forward and backward (streams.fork_join)
        enc2 = lambda enc: (lambda t: enc(t + enc(t)[0])[0])
        if x.is_cuda and x.dtype == torch.float32 and self.concurrent_encoders:
            x_l, x_a = fork_join(self, [enc2(self.encoder_l), enc2(self.encoder_a)], [x_l, x_a])
        else:
            x_l, x_a = enc2(self.encoder_l)(x_l), enc2(self.encoder_a)(x_a)
        x_l, x_a = x_l.permute(1, 0, 2), x_a.permute(1, 0, 2)
        qmask = qmask.to(x_l.dtype)
        h_f = self.dropout_rec(self.marn_cell_f(x, x_l, x_a, qmask))
        h_b = self.marn_cell_b(x, reverse_seq(x_l, umask), reverse_seq(x_a, umask), reverse_seq(qmask, umask))
        h_b = self.dropout_rec(reverse_seq(h_b, umask))
        h = torch.cat([h_f, h_b], dim=-1)
        attn1 = self.crossatt_l2a(x_l, x_a, self.w, self.v)          # CA2(w x_l, v x_a)          lsthm_sps.py:377-378
        attn2 = self.crossatt_a2l(x_a, x_l, self.v, self.w)
        attn1 = self.crossatt_l2a_1(x_a, attn1, self.v, self.v1)     # CA3(v x_a, v1 attn1)       lsthm_sps.py:382-383
        attn2 = self.crossatt_a2l_1(x_l, attn2, self.w, self.v2)
        output = self.fc[2](self.fc[1](linear3(torch.cat([h, attn1, attn2], dim=-1), self.fc[0].weight, self.fc[0].bias)))
        output = F.log_softmax(self.nn_out(output + x_l + x_a), 2).permute(1, 0, 2)
        return output.reshape(-1, output.size(-1)), x_l, x_a
