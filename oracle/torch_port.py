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


def perturb_ones(params_or_module, seed: int, std: float = 0.1):
    """The reference initialises every attention projection and fusion scalar of lsthm_sps to ones,
    which makes the model badly conditioned (SURVEY.md F6, §8c hazards).  Parity is therefore also
    checked with those tensors perturbed: in registration order, add N(0, std^2) drawn from one seeded
    generator to every all-ones tensor.  Works on a module or a name->tensor dict (same order)."""
    g = torch.Generator().manual_seed(seed)
    items = params_or_module.named_parameters() if hasattr(params_or_module, "named_parameters") \
        else params_or_module.items()
    with torch.no_grad():
        for _, t in items:
            if t.is_floating_point() and t.numel() > 0 and bool((t == 1).all()):
                t.add_((std * torch.randn(t.shape, generator=g)).to(t.device))
    return params_or_module


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
# lsthm_sps.py  (speaker-state variant, BASELINE.json configs[2])
# --------------------------------------------------------------------------------------
def lstm_cell(p: Params, pre: str, x, h, c):
    """torch.nn.LSTMCell semantics (gate order i,f,g,o; two bias vectors), used for the per-speaker
    cells lstm_q0 / lstm_q1 (model/lsthm_sps.py:146-147, 183, 188)."""
    g = F.linear(x, p[pre + ".weight_ih"], p[pre + ".bias_ih"]) + F.linear(h, p[pre + ".weight_hh"], p[pre + ".bias_hh"])
    d = h.shape[1]
    i, f, o = torch.sigmoid(g[:, :d]), torch.sigmoid(g[:, d:2 * d]), torch.sigmoid(g[:, 3 * d:])
    c2 = f * c + i * torch.tanh(g[:, 2 * d:3 * d])
    return o * torch.tanh(c2), c2


def lsthm1_cell(p: Params, pre: str, x, c_prev, h_prev, z_prev, s_prev):
    """LSTHM1.forward (model/lsthm_sps.py:28-44): LSTHM plus the speaker term S s; gates f,i,o,g."""
    s = _lin(p, pre + ".W", x) + _lin(p, pre + ".U", h_prev) + _lin(p, pre + ".V", z_prev) + _lin(p, pre + ".S", s_prev)
    dh = c_prev.shape[1]
    f, i, o = (torch.sigmoid(s[:, j * dh:(j + 1) * dh]) for j in range(3))
    g = torch.tanh(s[:, 3 * dh:])
    c = f * c_prev + i * g
    return c, torch.tanh(c) * o


def cross_attention_cell(p: Params, pre: str, x1, x2, tape, site):
    """In-cell CrossAttention.forward (model/lsthm_sps.py:59-72), executed the way the reference
    does (two outer products and a [N,128,128] product); Wv is never used."""
    dh = x1.shape[1]
    Q = torch.matmul(x1.unsqueeze(-1), p[pre + ".Wq"])           # [N, D, D]
    K = torch.matmul(x2.unsqueeze(-1), p[pre + ".Wk"])
    attn = torch.softmax(torch.matmul(Q / (dh ** 0.5), K), dim=-1)
    attn = _drop(attn, 0.2, site, tape)
    return torch.matmul(attn, x2.unsqueeze(-1)).squeeze(-1)


def party_rows(qmask_t: torch.Tensor):
    """Index form of MARN_cell._select_parties (model/lsthm_sps.py:238-259): ascending dialogue
    ids whose current speaker (argmax of the one-hot row; an all-zero padded row gives 0) is 0 / 1."""
    idx = torch.argmax(qmask_t, 1)
    return torch.nonzero(idx == 0).flatten(), torch.nonzero(idx == 1).flatten()


def sps_cell(p: Params, pre: str, x_l, x_a, qmask, tape: Optional[DropoutTape] = None):
    """MARN_cell.forward (model/lsthm_sps.py:156-221).  x_l,x_a [T,N,100], qmask [T,N,2] ->
    h [T,N,512] = [h_l | h_a | z_l | h_q].  NOTE the reference packs speaker-0 rows before
    speaker-1 rows and then treats packed row r as dialogue r; restated as is."""
    T, N, _ = x_l.shape
    dq = 128
    z = lambda: x_l.new_zeros(N, dq)
    h_l, h_a, h_q0, h_q1, c_l, c_a, c_q0, c_q1, z_l = (z() for _ in range(9))
    q = x_l.new_zeros(N, 2, dq)
    site = pre + ".dropout"
    out = []
    for t in range(T):
        P0, P1 = party_rows(qmask[t])
        N0, N1 = P0.numel(), P1.numel()
        q0_sel = torch.cat([q[P0, 0], x_l.new_zeros(N - N0, dq)], 0) if N0 else None
        q1_sel = torch.cat([q[P1, 1], x_l.new_zeros(N - N1, dq)], 0) if N1 else None
        if N0:
            h_q0, c_q0 = lstm_cell(p, pre + ".lstm_q0", q0_sel, h_q0, c_q0)
            h_q0 = _drop(h_q0, 0.5, site, tape)
        if N1:
            h_q1, c_q1 = lstm_cell(p, pre + ".lstm_q1", q1_sel, h_q1, c_q1)
            h_q1 = _drop(h_q1, 0.5, site, tape)
        if N0 and N1:
            h_q = torch.cat([h_q0[:N0], h_q1[:N1]], 0)
            h_0 = torch.cat([q0_sel[:N0], q1_sel[:N1]], 0)
        elif N0:
            h_q, h_0 = h_q0, q0_sel
        else:
            h_q, h_0 = h_q1, q1_sel
        m = qmask[t].unsqueeze(2)
        q = h_0.unsqueeze(1) * (1 - m) + h_q.unsqueeze(1) * m
        c_l, h_l = lsthm1_cell(p, pre + ".lsthm_l", x_l[t], c_l, h_l, z_l, h_q)
        h_l = _drop(h_l, 0.5, site, tape)
        c_a, h_a = lsthm1_cell(p, pre + ".lsthm_a", x_a[t], c_a, h_a, z_l, h_q)
        h_a = _drop(h_a, 0.5, site, tape)
        z_l = cross_attention_cell(p, pre + ".crossatt_l2a", c_l, c_a, tape, pre + ".crossatt_l2a.dropout")
        out.append(torch.cat([h_l, h_a, z_l, h_q], 1))
    return torch.stack(out, 0)


def reverse_seq(X: torch.Tensor, umask: torch.Tensor) -> torch.Tensor:
    """MARN1_sps._reverse_seq (model/lsthm_sps.py:396-410): flip each dialogue over its own length,
    zero-pad to the longest dialogue.  X [L,B,d], umask [B,L]."""
    lens = umask.sum(1).int().tolist()
    Lmax = max(lens)
    out = X.new_zeros(Lmax, X.shape[1], X.shape[2])
    for b, n in enumerate(lens):
        out[:n, b] = torch.flip(X[:n, b], [0])
    return out


def cross_attention_seq(p: Params, pre: str, x1, x2, tape, dk: int = 128):
    """CrossAttention2/3.forward (model/lsthm_sps.py:88-101, 116-129): dense, UNMASKED attention over
    the L utterances of each dialogue.  x1,x2 [L,B,*] -> [L,B,128]."""
    a, b = x1.permute(1, 0, 2), x2.permute(1, 0, 2)
    Q, K, V = torch.matmul(a, p[pre + ".Wq"]), torch.matmul(b, p[pre + ".Wk"]), torch.matmul(b, p[pre + ".Wv"])
    attn = torch.softmax(torch.matmul(Q / (dk ** 0.5), K.transpose(1, 2)), dim=-1)
    attn = _drop(attn, 0.2, pre + ".dropout", tape)
    return torch.matmul(attn, V).permute(1, 0, 2)


def sps_forward(p: Params, x, qmask, umask, tape: Optional[DropoutTape] = None, return_state: bool = False):
    """MARN1_sps.forward (model/lsthm_sps.py:349-394).  x [L,B,1124], qmask [L,B,2], umask [B,L] ->
    (log-probs [B*L, C] batch-major, x_l [L,B,100], x_a [L,B,100])."""
    xl = _lin(p, "linear_in", x[:, :, :1024].permute(1, 0, 2))
    xa = x[:, :, 1024:1124].permute(1, 0, 2)
    xl1 = encoder_layer(p, "encoder_l", xl, tape)
    xa1 = encoder_layer(p, "encoder_a", xa, tape)
    xl = encoder_layer(p, "encoder_l", xl + xl1, tape).permute(1, 0, 2)
    xa = encoder_layer(p, "encoder_a", xa + xa1, tape).permute(1, 0, 2)
    h_f = _drop(sps_cell(p, "marn_cell_f", xl, xa, qmask, tape), 0.5, "dropout_rec", tape)
    h_b = sps_cell(p, "marn_cell_b", reverse_seq(xl, umask), reverse_seq(xa, umask), reverse_seq(qmask, umask), tape)
    h_b = _drop(reverse_seq(h_b, umask), 0.5, "dropout_rec", tape)
    h = torch.cat([h_f, h_b], -1)
    w, v, v1, v2 = p["w"], p["v"], p["v1"], p["v2"]
    a1 = cross_attention_seq(p, "crossatt_l2a", w * xl, v * xa, tape)
    a2 = cross_attention_seq(p, "crossatt_a2l", v * xa, w * xl, tape)
    a1 = cross_attention_seq(p, "crossatt_l2a_1", v * xa, v1 * a1, tape)
    a2 = cross_attention_seq(p, "crossatt_a2l_1", w * xl, v2 * a2, tape)
    o = _drop(torch.relu(_lin(p, "fc.0", torch.cat([h, a1, a2], -1))), 0.5, "fc.2", tape)
    y = _drop(torch.relu(_lin(p, "nn_out.0", o + xl + xa)), 0.5, "nn_out.2", tape)
    logp = torch.log_softmax(_lin(p, "nn_out.3", y), 2).permute(1, 0, 2)
    logp = logp.reshape(-1, logp.shape[-1])
    if return_state:
        return logp, xl, xa, h
    return logp, xl, xa


# --------------------------------------------------------------------------------------
# lsthm_onlysp.py  (the reference's train.py default model, `--model MARN1_onlysp`)
# --------------------------------------------------------------------------------------
def gru_cell(p: Params, pre: str, x, h):
    """torch.nn.GRUCell semantics (gate order r,z,n; n = tanh(W_in x + b_in + r * (W_hn h + b_hn)))."""
    gi = F.linear(x, p[pre + ".weight_ih"], p[pre + ".bias_ih"])
    gh = F.linear(h, p[pre + ".weight_hh"], p[pre + ".bias_hh"])
    d = h.shape[1]
    r = torch.sigmoid(gi[:, :d] + gh[:, :d])
    z = torch.sigmoid(gi[:, d:2 * d] + gh[:, d:2 * d])
    n = torch.tanh(gi[:, 2 * d:] + r * gh[:, 2 * d:])
    return (1 - z) * n + z * h


def onlysp_cell(p: Params, pre: str, x_l, x_a, qmask, tape: Optional[DropoutTape] = None):
    """MARN_cell.forward of lsthm_onlysp (model/lsthm_onlysp.py:156-206): per-dialogue speaker state through one
    GRUCell on [x_l|x_a] and a DialogueRNN-style party update; LSTHM1 cells and the in-cell rank-1 attention as in
    lsthm_sps.  Dialogues are independent here (no packed rows).  Returns h [T,N,512] = [h_l|h_a|z_l|h_s]."""
    T, N, _ = x_l.shape
    z = lambda: x_l.new_zeros(N, 128)
    h_l, h_a, c_l, c_a, z_l = (z() for _ in range(5))
    q = x_l.new_zeros(N, 2, 128)
    site = pre + ".dropout"
    ar = torch.arange(N)
    out = []
    for t in range(T):
        U = torch.cat((x_l[t], x_a[t]), dim=1)
        qs_0 = q[ar, torch.argmax(qmask[t], 1)]                       # _select_parties, lines 200-205
        h_s = _drop(gru_cell(p, pre + ".gru_s", U, qs_0), 0.5, site, tape)
        m = qmask[t].unsqueeze(2)
        q = q * (1 - m) + h_s.unsqueeze(1) * m
        c_l, h_l = lsthm1_cell(p, pre + ".lsthm_l", x_l[t], c_l, h_l, z_l, h_s)
        h_l = _drop(h_l, 0.5, site, tape)
        c_a, h_a = lsthm1_cell(p, pre + ".lsthm_a", x_a[t], c_a, h_a, z_l, h_s)
        h_a = _drop(h_a, 0.5, site, tape)
        z_l = cross_attention_cell(p, pre + ".crossatt_l2a", c_l, c_a, tape, pre + ".crossatt_l2a.dropout")
        out.append(torch.cat([h_l, h_a, z_l, h_s], 1))
    return torch.stack(out, 0)


def onlysp_forward(p: Params, x, qmask, umask, tape: Optional[DropoutTape] = None):
    """MARN1_onlysp.forward (model/lsthm_onlysp.py:260-301): encoders applied twice WITHOUT residual (264-268),
    bidirectional cell, sequence cross-attention with the learnable scalars, head directly on the 1280-d concat."""
    xl = _lin(p, "linear_in", x[:, :, :1024].permute(1, 0, 2))
    xa = x[:, :, 1024:1124].permute(1, 0, 2)
    xl = encoder_layer(p, "encoder_l", xl, tape)
    xa = encoder_layer(p, "encoder_a", xa, tape)
    xl = encoder_layer(p, "encoder_l", xl, tape).permute(1, 0, 2)
    xa = encoder_layer(p, "encoder_a", xa, tape).permute(1, 0, 2)
    h_f = _drop(onlysp_cell(p, "marn_cell_f", xl, xa, qmask, tape), 0.5, "dropout_rec", tape)
    h_b = onlysp_cell(p, "marn_cell_b", reverse_seq(xl, umask), reverse_seq(xa, umask), reverse_seq(qmask, umask), tape)
    h_b = _drop(reverse_seq(h_b, umask), 0.5, "dropout_rec", tape)
    h = torch.cat([h_f, h_b], -1)
    w, v, v1, v2 = p["w"], p["v"], p["v1"], p["v2"]
    a1 = cross_attention_seq(p, "crossatt_l2a", w * xl, v * xa, tape)
    a2 = cross_attention_seq(p, "crossatt_a2l", v * xa, w * xl, tape)
    a1 = cross_attention_seq(p, "crossatt_l2a_1", v * xa, v1 * a1, tape)
    a2 = cross_attention_seq(p, "crossatt_a2l_1", w * xl, v2 * a2, tape)
    y = _drop(torch.relu(_lin(p, "nn_out.0", torch.cat([h, a1, a2], -1))), 0.5, "nn_out.2", tape)
    logp = torch.log_softmax(_lin(p, "nn_out.3", y), 2).permute(1, 0, 2)
    return logp.reshape(-1, logp.shape[-1]), xl, xa


# --------------------------------------------------------------------------------------
# lsthm_nsps.py  (speaker + listener party update, softmax(p) fusion weights, residual+LayerNorm cross attention)
# --------------------------------------------------------------------------------------
def nsps_cell(p: Params, pre: str, u, x_l, x_a, qmask, tape: Optional[DropoutTape] = None):
    """MARN_cell.forward of lsthm_nsps (model/lsthm_nsps.py:159-215): the GRU runs on the step's input features ``u``
    (line 176) and the current speaker's state; BOTH parties are then rewritten — the speaker's with the new state, the
    other with the *listener's previous* state (q_l, lines 178, 184-188).  Returns [T,N,512] = [h_l|h_a|z_l|h_s]
    (the reference returns the first 384 columns as ``h`` and h_l, h_a, h_s separately, lines 196-215)."""
    T, N, _ = x_l.shape
    z = lambda: x_l.new_zeros(N, 128)
    h_l, h_a, c_l, c_a, z_l = (z() for _ in range(5))
    q = x_l.new_zeros(N, 2, 128)
    site = pre + ".dropout"
    ar = torch.arange(N)
    out = []
    for t in range(T):
        idx = torch.argmax(qmask[t], 1)
        qs_0, ql_0 = q[ar, idx], q[ar, 1 - idx]                       # _select_parties, lines 232-239
        h_s = _drop(gru_cell(p, pre + ".gru_s", u[t], qs_0), 0.5, site, tape)
        m = qmask[t].unsqueeze(2)
        q = ql_0.unsqueeze(1) * (1 - m) + h_s.unsqueeze(1) * m
        c_l, h_l = lsthm1_cell(p, pre + ".lsthm_l", x_l[t], c_l, h_l, z_l, h_s)
        h_l = _drop(h_l, 0.5, site, tape)
        c_a, h_a = lsthm1_cell(p, pre + ".lsthm_a", x_a[t], c_a, h_a, z_l, h_s)
        h_a = _drop(h_a, 0.5, site, tape)
        z_l = cross_attention_cell(p, pre + ".crossatt_l2a", c_l, c_a, tape, pre + ".crossatt_l2a.dropout")
        out.append(torch.cat([h_l, h_a, z_l, h_s], 1))
    return torch.stack(out, 0)


def cross_attention_seq_ln(p: Params, pre: str, x1, x2, tape, dk: int = 100):
    """CrossAttention2.forward of lsthm_nsps (model/lsthm_nsps.py:90-108): the dense unmasked attention followed by
    ``LayerNorm(out + x1)`` (eps 1e-6)."""
    out = cross_attention_seq(p, pre, x1, x2, tape, dk) + x1
    return F.layer_norm(out, (out.shape[-1],), p[pre + ".layer_norm.weight"], p[pre + ".layer_norm.bias"], 1e-6)


def nsps_forward(p: Params, x, qmask, umask, tape: Optional[DropoutTape] = None, text_encoder: bool = True):
    """MARN1_nsps.forward (model/lsthm_nsps.py:300-359).  ``text_encoder=False`` is MARN1_no_en.forward
    (model/lsthm_no_en.py:300-359): the same function with the two ``encoder_l`` calls commented out (lines 306, 309), i.e. the
    text stream reaches the cells and the cross attention as ``linear_in``'s output."""
    xl0 = _lin(p, "linear_in", x[:, :, :1024].permute(1, 0, 2))
    xa0 = x[:, :, 1024:1124].permute(1, 0, 2)
    u = torch.cat([xl0, xa0], dim=2).permute(1, 0, 2)
    if text_encoder:
        xl1 = encoder_layer(p, "encoder_l", xl0, tape)
    xa1 = encoder_layer(p, "encoder_a", xa0, tape)
    xl = (encoder_layer(p, "encoder_l", xl0 + xl1, tape) if text_encoder else xl0).permute(1, 0, 2)
    xa = encoder_layer(p, "encoder_a", xa0 + xa1, tape).permute(1, 0, 2)
    rec = lambda t: _drop(t, 0.5, "dropout_rec", tape)
    o_f = nsps_cell(p, "marn_cell_f", u, xl, xa, qmask, tape)
    hf_l, hf_a, _ = rec(o_f[..., 0:128]), rec(o_f[..., 128:256]), rec(o_f[..., 384:512])
    o_b = nsps_cell(p, "marn_cell_b", reverse_seq(u, umask), reverse_seq(xl, umask), reverse_seq(xa, umask),
                    reverse_seq(qmask, umask), tape)
    o_b = reverse_seq(o_b, umask)
    rec(o_b[..., 0:384])                                              # h_b: dropped out but never used (lines 328, 335)
    hb_l, hb_a, _ = rec(o_b[..., 0:128]), rec(o_b[..., 128:256]), rec(o_b[..., 384:512])
    h_l, h_a = torch.cat([hf_l, hb_l], -1), torch.cat([hf_a, hb_a], -1)
    a1 = cross_attention_seq_ln(p, "crossatt_l2a", xl, xa, tape)
    a2 = cross_attention_seq_ln(p, "crossatt_a2l", xa, xl, tape)
    e = torch.exp(p["p"])
    w1, w2 = e[0] / e.sum(), e[1] / e.sum()
    resid_l = _drop(torch.relu(_lin(p, "fc.0", xl)), 0.5, "fc.2", tape)
    if tape is not None:
        _drop(torch.relu(_lin(p, "fc2.0", xa)), 0.5, "fc2.2", tape)   # resid_a: computed and discarded (line 352)
    fused = torch.cat([w1 * torch.cat([h_l, a2], 2), w2 * torch.cat([h_a, a1], 2)], dim=-1) + resid_l
    y = _drop(torch.relu(_lin(p, "nn_out.0", fused)), 0.5, "nn_out.2", tape)
    logp = torch.log_softmax(_lin(p, "nn_out.3", y), 2).permute(1, 0, 2)
    return logp.reshape(-1, logp.shape[-1]), xl, xa


def no_en_forward(p: Params, x, qmask, umask, tape: Optional[DropoutTape] = None):
    """MARN1_no_en.forward (model/lsthm_no_en.py:300-359): lsthm_nsps without the text encoder."""
    return nsps_forward(p, x, qmask, umask, tape, text_encoder=False)


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
