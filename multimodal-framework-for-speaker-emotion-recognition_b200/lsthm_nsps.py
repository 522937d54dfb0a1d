"""Drop-in for the reference's ``model/lsthm_nsps.py``: ``MARN1_nsps(n_classes, dataset)`` ->
``forward(x[L,B,1124], qmask[L,B,2], umask[B,L]) -> (log-probs [B*L, C], x_l, x_a)``.

This is the variant that holds the learnable-weight audio/text fusion in its softmax form (``w1, w2 = softmax(p)``,
lsthm_nsps.py:292, 347-355) and the residual + LayerNorm sequence cross attention (75-108).  Same constructor, parameter
names / shapes / registration order (including the never-used ``gru_l``, ``crossatt_a2l`` / ``Wv`` of the cell and ``fc2``
whose output the reference discards) and default-init RNG order as the reference.  ``MARN_cell.forward`` (159-215) runs as
the GRU speaker-state kernel pair with the listener party update (``lsthm_gsp_*``, listener = 1).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .encoder import EncoderLayer
from .lsthm_onlysp import MARN_cell as _GruCell
from .lsthm_sps import _SeqCrossAttention, reverse_seq
from .mm3 import linear3
from .streams import fork_join, state_without_streams


class CrossAttention2(_SeqCrossAttention):
    """lsthm_nsps.py:75-108: dense unmasked attention over the utterances, then residual + LayerNorm(eps 1e-6)."""

    def __init__(self, dh, dk, dv, attn_dropout=0.2):
        super().__init__(dh, dh, attn_dropout, dk=dk, dv=dv)
        self.layer_norm = nn.LayerNorm(dh, eps=1e-6)

    def forward(self, x_1, x_2):
        return self.layer_norm(super().forward(x_1, x_2) + x_1)


class MARN_cell(_GruCell):
    """lsthm_nsps.py:140-215: q[p] = q[listener party](1 - m_p) + h_s m_p (lines 184-188)."""
    listener = 1

    def _extra_cells(self):
        self.gru_s = nn.GRUCell(self.d_l + self.d_a, self.dh_s)
        self.gru_l = nn.GRUCell(self.d_l + self.d_a, self.dh_s)   # never used (line 156)


class MARN1_nsps(nn.Module):
    def __getstate__(self):
        return state_without_streams(self)          # cached CUDA streams are not part of the module's state

    text_encoder = True          # False in the MARN1_no_en variant (lsthm_no_en.py)

    def __init__(self, n_classes, dataset=None):
        super().__init__()
        self.d_l, self.d_a, self.d_r = 100, 100, 1024
        self.dh_l, self.dh_a, self.dh_sp, self.dh_li = 128, 128, 128, 128
        self.total_h_dim = self.dh_l + self.dh_a
        self.linear_in = nn.Linear(self.d_r, self.d_l)
        self.marn_cell_f = MARN_cell(self.dh_l, self.dh_a, self.d_l, self.d_a)
        self.marn_cell_b = MARN_cell(self.dh_l, self.dh_a, self.d_l, self.d_a)
        final_out = 2 * (self.total_h_dim + self.d_l)
        self.fc = nn.Sequential(nn.Linear(self.d_l, final_out), nn.ReLU(), nn.Dropout(0.5))
        self.fc2 = nn.Sequential(nn.Linear(self.d_a, final_out), nn.ReLU(), nn.Dropout(0.5))   # output discarded (line 352)
        self.nn_out = nn.Sequential(nn.Linear(final_out, 32), nn.ReLU(), nn.Dropout(0.5), nn.Linear(32, n_classes))
        self.dropout_rec = nn.Dropout(0.5)
        self.encoder_l = EncoderLayer(self.d_l, 40, 8, 40, 40)
        self.encoder_a = EncoderLayer(self.d_a, 40, 8, 40, 40)
        self.crossatt_l2a = CrossAttention2(self.d_l, self.d_l, self.d_l)
        self.crossatt_a2l = CrossAttention2(self.d_a, self.d_a, self.d_a)
        self.p = nn.Parameter(torch.ones(2))
        self.concurrent_encoders = True           # text / audio encoder chains on two CUDA streams (same results bit for bit)

    def forward(self, x, qmask, umask):
        x_l = linear3(x[:, :, :self.d_r].permute(1, 0, 2), self.linear_in.weight, self.linear_in.bias)
        x_a = x[:, :, self.d_r:self.d_r + self.d_a].permute(1, 0, 2)
        u = torch.cat([x_l, x_a], dim=2).permute(1, 0, 2)        # GRU input: PRE-encoder features (line 306)
        enc2 = lambda enc: (lambda t: enc(t + enc(t)[0])[0])    # enc(x + enc(x)), lsthm_nsps.py:306-310
        if not self.text_encoder:                                # lsthm_no_en.py:306,309 comments the two encoder_l calls out
            x_a = enc2(self.encoder_a)(x_a)
        elif x.is_cuda and x.dtype == torch.float32 and self.concurrent_encoders:
            x_l, x_a = fork_join(self, [enc2(self.encoder_l), enc2(self.encoder_a)], [x_l, x_a])   # two streams (streams.py)
        else:
            x_l, x_a = enc2(self.encoder_l)(x_l), enc2(self.encoder_a)(x_a)
        x_l, x_a = x_l.permute(1, 0, 2), x_a.permute(1, 0, 2)
        qmask = qmask.to(x_l.dtype)
        drop = self.dropout_rec
        o_f = self.marn_cell_f(u, x_l, x_a, qmask)               # [L,B,512] = [h_l | h_a | z_l | h_s]
        hf_l, hf_a = drop(o_f[..., 0:128]), drop(o_f[..., 128:256])
        drop(o_f[..., 384:512])                                   # hf_sp: drawn (RNG order) but unused downstream (line 319)
        o_b = self.marn_cell_b(reverse_seq(u, umask), reverse_seq(x_l, umask), reverse_seq(x_a, umask), reverse_seq(qmask, umask))
        o_b = reverse_seq(o_b, umask)
        drop(o_b[..., 0:384])                                     # h_b: likewise only consumes the RNG (line 328)
        hb_l, hb_a = drop(o_b[..., 0:128]), drop(o_b[..., 128:256])
        drop(o_b[..., 384:512])
        h_l, h_a = torch.cat([hf_l, hb_l], dim=-1), torch.cat([hf_a, hb_a], dim=-1)
        attn1 = self.crossatt_l2a(x_l, x_a)
        attn2 = self.crossatt_a2l(x_a, x_l)
        w = torch.softmax(self.p, 0)                              # exp(p_i) / sum exp(p) (lines 347-348)
        resid_l = self.fc[2](linear3(x_l, self.fc[0].weight, self.fc[0].bias, relu=True))
        fused = torch.cat([w[0] * h_l, w[0] * attn2, w[1] * h_a, w[1] * attn1], dim=-1) + resid_l
        y = self.nn_out[2](self.nn_out[1](linear3(fused, self.nn_out[0].weight, self.nn_out[0].bias)))
        output = F.log_softmax(linear3(y, self.nn_out[3].weight, self.nn_out[3].bias), 2).permute(1, 0, 2)
        return output.reshape(-1, output.size(-1)), x_l, x_a
