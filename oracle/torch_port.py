"""TEST INFRASTRUCTURE — CPU restatement (PyTorch ops, autograd) of the reference hot path.

This file is the *oracle*: a from-scratch functional restatement of what the reference
computes for the LSTHM hybrid recurrence and the blocks either side of it.  It is used
only as a checker (tests/, ``__graft_entry__.smoke()``) and as the ``cpu_baseline`` /
``--impl reference`` arm of bench.py.  The product path (the CUDA kernels behind the
C-ABI) never imports it.

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against *outputs of the reference itself*: ``oracle/make_golden.py``
imports /root/reference in the build container, runs it on seeded inputs and commits
the results under tests/golden/; ``tests/test_oracle_golden.py`` checks this file
against those fixtures (and, when /root/reference is present, against the live code).

Every function cites the reference lines it restates.  All functions take the model's
``state_dict`` (names identical to the reference's) so that no ``nn.Module`` of ours is
involved in the check.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# Dropout mask tape: the reference applies nn.Dropout at many sites,
# including on recurrent state, so train-mode parity needs both sides to consume the
# same masks.  A tape maps (site name, call index) -> mask already scaled by 1/(1-p).
# --------------------------------------------------------------------------------------
class DropoutTape:
    def __init__(self, seed: int = 0):
        self.masks: Dict[str, list] = {}
        self._cursor: Dict[str, int] = {}
        self._gen = torch.Generator().manual_seed(seed)
        self.recording = True

    def rewind(self) -> "DropoutTape":
        self._cursor = {}
        self.recording = False
        return self

    def mask(self, site: str, shape, p: float, dtype=torch.float32) -> torch.Tensor:
        i = self._cursor.get(site, 0)
        self._cursor[site] = i + 1
        lst = self.masks.setdefault(site, [])
        if i < len(lst):
            m = lst[i]
            assert tuple(m.shape) == tuple(shape), (site, i, m.shape, shape)
            return m.to(dtype)
        assert self.recording, f"tape exhausted at {site}[{i}]"
        keep = torch.bernoulli(torch.full(tuple(shape), 1.0 - p), generator=self._gen)
        m = keep / (1.0 - p)
        lst.append(m)
        return m.to(dtype)

    def stacked(self, site: str) -> torch.Tensor:
        return torch.stack(self.masks[site], 0)


def _drop(x: torch.Tensor, p: float, site: str, tape: Optional[DropoutTape]):
    """nn.Dropout at a named site: identity without a tape (== eval mode)."""
    if tape is None or p == 0.0:
        return x
    return x * tape.mask(site, x.shape, p, x.dtype)


def _lin(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, p[name + ".weight"], p.get(name + ".bias"))


# --------------------------------------------------------------------------------------
# encoder.py
# --------------------------------------------------------------------------------------
def encoder_layer(p: Params, pre: str, x: torch.Tensor, tape: Optional[DropoutTape] = None,
                  n_head: int = 8, d_k: int = 40) -> torch.Tensor:
    """One post-LN transformer layer over the L utterances of each dialogue, *no mask*.

    Restates EncoderLayer.forward (model/encoder.py:129-133) =
    MultiHeadAttention.forward (27-60) with ScaledDotProductAttention (71-86, temperature
    sqrt(d_k), dropout on the attention weights) followed by PositionwiseFeedForward
    (101-113; its ``fc`` is never applied).  x: [B, L, d].
    """
    B, L, d = x.shape
    a = pre + ".slf_attn."
    q = F.linear(x, p[a + "w_qs.weight"]).view(B, L, n_head, d_k).transpose(1, 2)
    k = F.linear(x, p[a + "w_ks.weight"]).view(B, L, n_head, d_k).transpose(1, 2)
    v = F.linear(x, p[a + "w_vs.weight"]).view(B, L, n_head, d_k).transpose(1, 2)
    w = torch.softmax(torch.matmul(q / math.sqrt(d_k), k.transpose(2, 3)), dim=-1)
    w = _drop(w, 0.1, a + "attention.dropout", tape)
    ctx = torch.matmul(w, v).transpose(1, 2).reshape(B, L, n_head * d_k)
    y = _drop(F.linear(ctx, p[a + "fc.weight"]), 0.1, a + "dropout", tape) + x
    y = F.layer_norm(y, (d,), p[a + "layer_norm.weight"], p[a + "layer_norm.bias"], 1e-6)
    f = pre + ".pos_ffn."
    h = F.linear(torch.relu(F.linear(y, p[f + "w_1.weight"], p[f + "w_1.bias"])),
                 p[f + "w_2.weight"], p[f + "w_2.bias"])
    h = _drop(h, 0.1, f + "dropout", tape) + y
    return F.layer_norm(h, (d,), p[f + "layer_norm.weight"], p[f + "layer_norm.bias"], 1e-6)


# --------------------------------------------------------------------------------------
# HybridRNN_AT.py / HybridRNN_ATV.py
# --------------------------------------------------------------------------------------
MAB_SPECS = {
    # name: (modalities, feature dims, cell sizes)   HybridRNN_ATV.py:43-45 / HybridRNN_AT.py:43-45
    "ATV": (("l", "a", "v"), (100, 100, 512), (128, 16, 64)),
    "AT": (("l", "a"), (100, 100), (128, 16)),
}


def lsthm_cell(p: Params, pre: str, x, c_prev, h_prev, z_prev):
    """LSTHM.forward (model/HybridRNN_ATV.py:21-37): gates f,i,o,g from W x + U h + V z."""
    s = _lin(p, pre + ".W", x) + _lin(p, pre + ".U", h_prev) + _lin(p, pre + ".V", z_prev)
    dh = c_prev.shape[1]
    f, i, o = (torch.sigmoid(s[:, j * dh:(j + 1) * dh]) for j in range(3))
    g = torch.tanh(s[:, 3 * dh:])
    c = f * c_prev + i * g
    return c, torch.tanh(c) * o


def mab_step(p: Params, mods, dhs, cs: torch.Tensor, tape, n_att: int = 4) -> torch.Tensor:
    """Multi-attention block + fc of one step (model/HybridRNN_ATV.py:123-129).

    cs: [N, D] concatenated new cell states.  Returns z_t [N, D].  The reference stacks
    the 4 heads on the batch axis and softmaxes over the D features; that is the same as
    viewing att(cs) as [N, 4, D] and softmaxing the last axis.
    """
    N, D = cs.shape
    a = torch.softmax(_lin(p, "att.0", cs).view(N, n_att, D), dim=-1)
    attended = a * cs.unsqueeze(1)                                  # [N, 4, D]
    red, o = [], 0
    for m, dh in zip(mods, dhs):
        red.append(_lin(p, f"reduce_dim_nn_{m}.0", attended[:, :, o:o + dh].reshape(N, n_att * dh)))
        o += dh
    u = torch.relu(_lin(p, "fc.0", torch.cat(red, 1)))
    u = _drop(u, 0.3, "fc.2", tape)
    return _lin(p, "fc.3", u)


def mab_forward(p: Params, x: torch.Tensor, kind: str = "ATV", tape: Optional[DropoutTape] = None,
                return_state: bool = False):
    """MARN.forward of HybridRNN_AT/ATV (model/HybridRNN_ATV.py:84-155).

    x: [T, N, sum(d)] time-major.  Returns probabilities [T*N, C] time-major (line 153).
    Per-step structure is kept exactly as the reference executes it (W·x inside the loop,
    per-step head) so that timing this function is a fair stand-in for the reference.
    """
    mods, ds, dhs = MAB_SPECS[kind]
    T, N, _ = x.shape
    D = sum(dhs)
    xs, o = [], 0
    for m, d in zip(mods, ds):
        xm = encoder_layer(p, f"encoder_{m}", x[:, :, o:o + d].permute(1, 0, 2), tape)
        xs.append(xm.permute(1, 0, 2))
        o += d
    c = [x.new_zeros(N, dh) for dh in dhs]
    h = [x.new_zeros(N, dh) for dh in dhs]
    z = x.new_zeros(N, D)
    outs, hz = [], []
    for t in range(T):
        new = [lsthm_cell(p, f"lsthm_{m}", xs[k][t], c[k], h[k], z) for k, m in enumerate(mods)]
        c = [cn for cn, _ in new]
        h = [hn for _, hn in new]
        z = mab_step(p, mods, dhs, torch.cat(c, 1), tape)
        all_hs = torch.cat(h + [z], 1)
        hz.append(all_hs)
        y = torch.relu(_lin(p, "nn_out.0", all_hs))                 # Dropout(0.0) is a no-op
        outs.append(torch.softmax(_lin(p, "nn_out.3", y), dim=-1))
    out = torch.cat(outs, 0)
    if return_state:
        return out, torch.stack(hz, 0), xs
    return out


def mab_gate_inputs(p: Params, xs, kind: str = "ATV") -> torch.Tensor:
    """Hoisted input projection: gx[t] = cat_m(W_m x_m[t] + bW_m + bU_m + bV_m)  [T, N, 4D].
    Same sums as LSTHM.forward lines 23-27 with the three biases gathered."""
    mods = MAB_SPECS[kind][0]
    return torch.cat([F.linear(xs[k], p[f"lsthm_{m}.W.weight"],
                               p[f"lsthm_{m}.W.bias"] + p[f"lsthm_{m}.U.bias"] + p[f"lsthm_{m}.V.bias"])
                      for k, m in enumerate(mods)], dim=-1)


def mab_recurrence(p: Params, gx: torch.Tensor, kind: str = "ATV", tape: Optional[DropoutTape] = None):
    """The recurrence of HybridRNN_ATV.py:117-143 at the CUDA kernels' boundary:
    gx [T, N, 4D] (see mab_gate_inputs) -> hz [T, N, 2D] = [h_t | z_t]."""
    mods, _, dhs = MAB_SPECS[kind]
    T, N, _ = gx.shape
    D = sum(dhs)
    c = [gx.new_zeros(N, dh) for dh in dhs]
    h = [gx.new_zeros(N, dh) for dh in dhs]
    z = gx.new_zeros(N, D)
    hz = []
    for t in range(T):
        go = 0
        for k, (m, dh) in enumerate(zip(mods, dhs)):
            s = gx[t, :, go:go + 4 * dh] + F.linear(h[k], p[f"lsthm_{m}.U.weight"]) \
                + F.linear(z, p[f"lsthm_{m}.V.weight"])
            f, i, o = (torch.sigmoid(s[:, j * dh:(j + 1) * dh]) for j in range(3))
            c[k] = f * c[k] + i * torch.tanh(s[:, 3 * dh:])
            h[k] = torch.tanh(c[k]) * o
            go += 4 * dh
        z = mab_step(p, mods, dhs, torch.cat(c, 1), tape)
        hz.append(torch.cat(h + [z], 1))
    return torch.stack(hz, 0)


def mab_head(p: Params, hz: torch.Tensor) -> torch.Tensor:
    """nn_out applied to every step's [h|z] (HybridRNN_ATV.py:68-73,139-141,153): [T,N,2D] -> [T*N, C]."""
    y = torch.relu(_lin(p, "nn_out.0", hz.reshape(-1, hz.shape[-1])))
    return torch.softmax(_lin(p, "nn_out.3", y), dim=-1)


# --------------------------------------------------------------------------------------
# loss.py
# --------------------------------------------------------------------------------------
def masked_loss(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor, kind: str = "ce"):
    """MaskedLoss.forward with weight=None (loss.py:13-21): sum-reduced loss over pred*mask,
    divided by the number of real utterances.  ``kind`` = 'ce' (CrossEntropyLoss, the
    train.py default) or 'nll'."""
    pm = pred * mask.reshape(-1, 1)
    if kind == "ce":
        s = F.cross_entropy(pm, target, reduction="sum")
    else:
        s = F.nll_loss(pm, target, reduction="sum")
    return s / mask.sum()
