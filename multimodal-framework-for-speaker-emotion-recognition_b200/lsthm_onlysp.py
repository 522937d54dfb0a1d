"""Drop-in for the reference's ``model/lsthm_onlysp.py`` (the default model of train.py):
``MARN1_onlysp(n_classes)`` -> ``forward(x[L,B,1124], qmask[L,B,2], umask[B,L]) -> (log-probs [B*L, C], x_l, x_a)``.

Same constructor, parameter names / shapes / registration order (including the never-used ones: ``lstm_q0``, ``lstm_q1``,
``lstm_s``, ``crossatt_a2l`` and ``crossatt_*.Wv`` of the cell, ``linear`` of the model) and default-init RNG order as the
reference (lsthm_onlysp.py:9-20, 46-56, 73-83, 101-111, 129-154, 213-258).  ``MARN_cell.forward`` (156-198) runs as one
fused CUDA kernel pair per direction (``lsthm_gsp_*``, csrc/sps_kernels.cuh MODE 1); encoders, heads and the hoisted input
projections are the shared time-parallel kernels; the sequence-level cross attention is shared with lsthm_sps.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .encoder import EncoderLayer
from .gsp_recurrence import gsp_cell
from .lsthm_sps import LSTHM1, CrossAttention, CrossAttention2, CrossAttention3, reverse_seq
from .streams import fork_join, state_without_streams
from .mm3 import linear3, linear_cat


class MARN_cell(nn.Module):
    """lsthm_onlysp.py:129-198.  ``listener`` selects the party update of lsthm_nsps (see lsthm_nsps.py here)."""
    listener = 0

    def __init__(self, dh_l, dh_a, d_l, d_a, dropout=0.5):
        super().__init__()
        self.crossatt_l2a = CrossAttention()
        self.crossatt_a2l = CrossAttention()          # never used by the reference either (line 192)
        self.dh_l, self.dh_a, self.dh_q, self.d_l, self.d_a, self.dh_s = dh_l, dh_a, dh_l, d_l, d_a, 128
        self.speaker_size = 4 * dh_l
        self.lsthm_l = LSTHM1(dh_l, d_l, dh_l, self.dh_s)
        self.lsthm_a = LSTHM1(dh_a, d_a, dh_l, self.dh_s)
        self._extra_cells()
        self.dropout = nn.Dropout(dropout)
        self.rows_per_cta = 0
        self.mask_override = None      # test hook: (ms, ml, ma, att_mask) dropout mask tape

    def _extra_cells(self):
        self.lstm_q0 = nn.LSTMCell(self.dh_s, self.dh_s)   # never used (lines 149-150, 153)
        self.lstm_q1 = nn.LSTMCell(self.dh_s, self.dh_s)
        self.gru_s = nn.GRUCell(self.d_l + self.d_a, self.dh_s)
        self.lstm_s = nn.LSTMCell(self.dh_s, self.dh_s)

    def cell_weights(self):
        l, a, g = self.lsthm_l, self.lsthm_a, self.gru_s
        return [l.U.weight, a.U.weight, l.V.weight, a.V.weight, l.S.weight, a.S.weight, g.weight_hh, g.bias_hh,
                self.crossatt_l2a.Wq, self.crossatt_l2a.Wk]

    def forward(self, u, x_l, x_a, qmask):
        """u [T,N,200]: the GRU input of every step (onlysp: cat[x_l, x_a], line 173; nsps: the pre-encoder features)."""
        T, N, _ = x_l.shape
        l, a = self.lsthm_l, self.lsthm_a
        gx = linear_cat([x_l, x_a], [l.W.weight, a.W.weight],
                        [l.W.bias + l.U.bias + l.V.bias + l.S.bias, a.W.bias + a.U.bias + a.V.bias + a.S.bias]).view(T, N, 2, 512)
        gxs = linear3(u, self.gru_s.weight_ih, self.gru_s.bias_ih)                               # [T,N,384]
        att_p, seed = 0.0, 0
        if self.mask_override is not None:
            masks = self.mask_override
        elif self.training:
            p = self.dropout.p
            draw = lambda: torch.empty(T, N, 128, device=gx.device).bernoulli_(1 - p).mul_(1 / (1 - p)) if p > 0 else None
            masks = (draw(), draw(), draw(), None)
            att_p = self.crossatt_l2a.dropout.p
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        else:
            masks = (None,) * 4
        return gsp_cell(gx, gxs, qmask, masks, self.cell_weights(), self.listener, self.rows_per_cta, att_p, seed)


class MARN1_onlysp(nn.Module):
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
        self.num_atts = 4
        final_out = 2 * (self.total_h_dim + self.dh_l + self.dh_l) + self.dh_l + self.dh_a
        self.linear = nn.Linear(final_out, 32)            # never used (line 231)
        self.nn_out = nn.Sequential(nn.Linear(final_out, 32), nn.ReLU(), nn.Dropout(0.5), nn.Linear(32, n_classes))
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
        # the encoders are applied twice, without residual (264-268); text and audio are independent up to the cells: on a CUDA
        # device the two chains run on two streams, forward and backward (streams.fork_join)
        enc2 = lambda enc: (lambda t: enc(enc(t)[0])[0])
        if x.is_cuda and x.dtype == torch.float32 and self.concurrent_encoders:
            x_l, x_a = fork_join(self, [enc2(self.encoder_l), enc2(self.encoder_a)], [x_l, x_a])
        else:
            x_l, x_a = enc2(self.encoder_l)(x_l), enc2(self.encoder_a)(x_a)
        x_l, x_a = x_l.permute(1, 0, 2), x_a.permute(1, 0, 2)
        qmask = qmask.to(x_l.dtype)
        h_f = self.dropout_rec(self.marn_cell_f(torch.cat([x_l, x_a], -1), x_l, x_a, qmask))
        r_l, r_a = reverse_seq(x_l, umask), reverse_seq(x_a, umask)
        h_b = self.marn_cell_b(torch.cat([r_l, r_a], -1), r_l, r_a, reverse_seq(qmask, umask))
        h_b = self.dropout_rec(reverse_seq(h_b, umask))
        h = torch.cat([h_f, h_b], dim=-1)
        attn1 = self.crossatt_l2a(x_l, x_a, self.w, self.v)          # CA2(w x_l, v x_a)          lsthm_onlysp.py:287-288
        attn2 = self.crossatt_a2l(x_a, x_l, self.v, self.w)
        attn1 = self.crossatt_l2a_1(x_a, attn1, self.v, self.v1)     # CA3(v x_a, v1 attn1)       lsthm_onlysp.py:292-293
        attn2 = self.crossatt_a2l_1(x_l, attn2, self.w, self.v2)
        y = self.nn_out[2](self.nn_out[1](linear3(torch.cat([h, attn1, attn2], dim=-1), self.nn_out[0].weight, self.nn_out[0].bias)))
        output = F.log_softmax(linear3(y, self.nn_out[3].weight, self.nn_out[3].bias), 2).permute(1, 0, 2)
        return output.reshape(-1, output.size(-1)), x_l, x_a
