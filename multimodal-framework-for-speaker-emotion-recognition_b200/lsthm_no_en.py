"""Drop-in for the reference's ``model/lsthm_no_en.py``: ``MARN1_no_en(n_classes, dataset)`` ->
``forward(x[L,B,1124], qmask[L,B,2], umask[B,L]) -> (log-probs [B*L, C], x_l, x_a)``.

The reference file is ``lsthm_nsps.py`` with the two ``encoder_l`` calls of ``forward`` commented out (lsthm_no_en.py:306,
309): the text stream reaches the cells, the sequence cross attention and the residual branch as ``linear_in``'s output.
``encoder_l`` is still constructed (line 287), so it is registered here as well — same state_dict and default-init RNG order —
and, as in the reference step, never receives a gradient (``ddp.unused_parameter_names`` lists it).
"""
from __future__ import annotations

from .lsthm_nsps import MARN1_nsps, MARN_cell, CrossAttention2  # noqa: F401  (same cell and attention classes, lines 75-215)


class MARN1_no_en(MARN1_nsps):
    text_encoder = False
