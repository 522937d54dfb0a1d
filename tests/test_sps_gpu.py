"""GPU suite for the speaker-state cell (lsthm_sps, BASELINE.json configs[2]) — all through the C ABI.

  1. MARN_cell level: kernel (fwd + BPTT) vs the oracle's torch restatement + autograd, on random
     two-speaker dialogues incl. steps where one party is absent (the skipped-LSTMCell branch), ragged
     tiles, several tile heights, eval and full mask tape (recurrent-state dropout + attention dropout);
  2. MARN1_sps module vs the reference-generated fixtures (eval, perturbed ones-parameters, train tape);
  3. properties at N=1024, T=110: determinism, tile-height invariance, inference == training forward,
     in-kernel attention dropout is seeded and active;
  4. the documented cross-dialogue coupling of the reference is reproduced, i.e. the
     kernel does NOT treat dialogues as independent.
"""
from importlib import import_module

import numpy as np
import pytest
import torch

import lsthm_b200
from helpers import (sps_port_run, sps_port_run64, sps_kink_truths, TOL_GRAD, TOL_OUT, check_against_fp64_truth, check_against_golden, e_inf, golden_files, load_golden, sps_cell_masks,
                     sps_run_module, sps_seeded_model)
from oracle import torch_port as tp

pytestmark = pytest.mark.gpu


def _dialogues(T, N, g, absent_step=None):
    q = torch.zeros(T, N, 2)
    s = torch.randint(0, 2, (N,), generator=g)
    for t in range(T):
        flip = torch.rand(N, generator=g) < 0.6
        s = torch.where(flip, 1 - s, s)
        q[t, torch.arange(N), s] = 1
    if absent_step is not None:          # nobody is speaker 1 at this step / nobody is speaker 0 at the next
        q[absent_step] = torch.tensor([1.0, 0.0])
        if absent_step + 1 < T:
            q[absent_step + 1] = torch.tensor([0.0, 1.0])
    return q


@pytest.mark.parametrize("T,N,rows,masked", [(6, 5, 4, False), (5, 11, 4, True), (7, 9, 8, False), (4, 13, 7, True),
                                             (6, 64, 7, False), (5, 64, 7, True),      # production tile height, 10 CTAs, ragged last tile
                                             (3, 3, 1, False), (5, 6, 2, True), (4, 10, 3, False), (3, 16, 5, False)])
def test_cell_vs_oracle(T, N, rows, masked):
    g = torch.Generator().manual_seed(T * 100 + N)
    model = sps_seeded_model(200 + N, True)
    cell = model.marn_cell_f
    pre = "marn_cell_f"
    x_l, x_a = torch.randn(T, N, 100, generator=g), torch.randn(T, N, 100, generator=g)
    qmask = _dialogues(T, N, g, absent_step=1 if T > 3 else None)
    dout = torch.randn(T, N, 512, generator=g)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    xl_r, xa_r = x_l.clone().requires_grad_(True), x_a.clone().requires_grad_(True)
    tape = tp.DropoutTape(7) if masked else None
    ref = tp.sps_cell(params, pre, xl_r, xa_r, qmask, tape)
    (ref * dout).sum().backward()
    model = model.to("cuda").eval()          # dropout comes only from the explicit mask tape below
    cell.rows_per_cta = rows
    if masked:
        cell.mask_override = sps_cell_masks(tape, pre, qmask, "cuda")
    xl_c, xa_c = x_l.cuda().requires_grad_(True), x_a.cuda().requires_grad_(True)
    out = cell(None, xl_c, xa_c, qmask.cuda())
    (out * dout.cuda()).sum().backward()
    errs = {"out": e_inf(out.detach().cpu(), ref.detach()), "dx_l": e_inf(xl_c.grad.cpu(), xl_r.grad),
            "dx_a": e_inf(xa_c.grad.cpu(), xa_r.grad)}
    for n, p in model.named_parameters():
        if n.startswith(pre + ".") and params[n].grad is not None:
            assert p.grad is not None, n
            errs[n] = e_inf(p.grad.cpu(), params[n].grad)
        elif n.startswith(pre + "."):
            assert p.grad is None, n
    assert errs["out"] <= 2e-5, errs
    assert max(errs.values()) <= 2e-4, sorted(errs.items(), key=lambda kv: -kv[1])[:6]


@pytest.mark.parametrize("path", golden_files("sps_*.npz"), ids=lambda p: p.split("/")[-1][:-4])
def test_module_matches_reference_fixture(path):
    fix = load_golden(path)
    logp, loss, dx, grads = sps_run_module(fix)
    # the fp32 reference itself is 1e-5..1.4e-3 away from its own fp64 run on these (ill-conditioned) cases,
    # so the bar is relative to the fp64 truth: err <= max(1e-4 | 1e-3, 3 x reference-fp32 error)
    try:
        errs = check_against_fp64_truth(fix, logp, loss, dx, grads, tol_out=TOL_OUT, tol_grad=TOL_GRAD)
    except AssertionError as first:
        # A ReLU pre-activation closer to zero than the forward tolerance (1e-4 of the layer's scale) has no defined
        # derivative at parity resolution: an implementation that is 1e-5 away in the forward may legitimately sit on
        # the other side of the kink (observed: fc.0 unit 59 of one utterance in sps_s111, whose flip moves dx by
        # 1.8e-3 although two different fp32-exact attention implementations already disagree on it).  Accept the run
        # iff it meets the SAME bars against the fp64 truth with some subset of those near-kink units flipped.
        if int(fix["train"]):
            raise
        errs = None
        for sub, truth in sps_kink_truths(fix):
            try:
                errs = check_against_fp64_truth(fix, logp, loss, dx, grads, tol_out=TOL_OUT, tol_grad=TOL_GRAD, truth=truth)
                errs["kink_flips"] = [(site, idx) for _, site, idx, _ in sub]
                break
            except AssertionError:
                continue
        if errs is None:
            raise first
        # the escape hatch is pinned: only the un-perturbed s111 fixture has a unit that close to its kink, and it is ONE unit
        # (ReLU call 4 of the oracle's forward = fc.0 of the head, flat index 2459 = utterance 24, unit 59)
        assert path.endswith("sps_s111_T9_N5_eval.npz") and errs["kink_flips"] == [(4, 2459)], (path, errs["kink_flips"])
        print("KINK PATH TAKEN:", path.split("/")[-1], errs["kink_flips"])
    if not int(fix["perturb"]) and "kink_flips" not in errs:
        check_against_golden(fix, logp, loss, dx, grads, tol_out=TOL_OUT, tol_grad=TOL_GRAD)
    print(path.split("/")[-1], errs)


def _batch(T, N, seed, full=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, N, 1124, generator=g)
    qmask = _dialogues(T, N, g)
    umask = torch.ones(N, T)
    labels = torch.randint(0, 6, (N * T,), generator=g)
    return x.cuda(), qmask.cuda(), umask.cuda(), labels.cuda()


def _fwd_bwd(model, x, qmask, umask, labels):
    model.zero_grad(set_to_none=True)
    xx = x.clone().requires_grad_(True)
    logp, _, _ = model(xx, qmask, umask)
    tp.masked_loss(logp, labels, umask, "ce").backward()
    return logp.detach(), xx.grad.detach()


def test_properties_at_benchmark_size():
    T, N = 110, 1024
    model = sps_seeded_model(111, True, "cuda").eval()
    x, qmask, umask, labels = _batch(T, N, 3)
    l0, d0 = _fwd_bwd(model, x, qmask, umask, labels)
    assert torch.isfinite(l0).all() and torch.isfinite(d0).all()
    l1, d1 = _fwd_bwd(model, x, qmask, umask, labels)
    assert torch.equal(l0, l1) and torch.equal(d0, d1)                      # deterministic
    model.marn_cell_f.rows_per_cta = model.marn_cell_b.rows_per_cta = 8
    l8, d8 = _fwd_bwd(model, x, qmask, umask, labels)
    model.marn_cell_f.rows_per_cta = model.marn_cell_b.rows_per_cta = 0
    assert torch.equal(l0, l8) and torch.equal(d0, d8)                      # tile height does not change the math
    with torch.no_grad():
        li, _, _ = model(x, qmask, umask)
    assert torch.equal(li, l0)                                              # inference path == training-path forward


def test_module_vs_oracle_multi_cta_shard():
    """The whole module on a shard that spans several CTAs at the production tile height (N = 64, 7 rows per CTA: 10 CTAs
    with a ragged last tile, both grid exchanges live) against the oracle's restatement.  The reference couples the
    dialogues of a shard through the packed rows, so the whole shard is compared.  Forward quantities only:
    with 84 000 ReLU units in the shard a handful always sit within the forward tolerance of their kink, where gradients
    are undefined at parity resolution — the adjoints of the multi-CTA path are checked unit-free at the
    cell level (test_cell_vs_oracle, N = 64 cases) and on the fixtures."""
    import numpy as np
    T, N, seed = 10, 64, 123
    g = torch.Generator().manual_seed(seed)
    fix = {"seed": np.array(seed), "perturb": np.array(1), "train": np.array(0), "sample_stride": np.array(97),
           "x": torch.randn(T, N, 1124, generator=g).numpy(), "qmask": _dialogues(T, N, g, absent_step=3).numpy(),
           "umask": np.ones((N, T), np.float32), "labels": torch.randint(0, 6, (N, T), generator=g).numpy()}
    lp32, loss32, _, _ = sps_port_run(fix)
    truth, _ = sps_port_run64(fix)
    logp, loss, dx, grads = sps_run_module(fix, rows_per_cta=7)
    bar = lambda tol, a32, a64: max(tol, 3.0 * e_inf(a32, a64))
    assert e_inf(logp, truth["probs64"]) <= bar(TOL_OUT, lp32, truth["probs64"]), e_inf(logp, truth["probs64"])
    assert abs(float(loss) - float(truth["loss64"])) / abs(float(truth["loss64"])) <= TOL_OUT
    top2 = np.sort(truth["probs64"], -1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 1e-3
    assert (np.argmax(np.asarray(logp), -1)[decided] == np.argmax(truth["probs64"], -1)[decided]).all()
    assert torch.isfinite(dx).all() and all(torch.isfinite(v).all() for v in grads.values() if v is not None)


def test_reference_coupling_is_reproduced():
    """Dialogue 0 alone vs inside a batch differ in the reference (packed rows leak speaker state across
    dialogues); our kernel must show the same dependence, and agree with the oracle on both."""
    T, N = 8, 4
    model = sps_seeded_model(5, True)
    g = torch.Generator().manual_seed(1)
    x_l, x_a = torch.randn(T, N, 100, generator=g), torch.randn(T, N, 100, generator=g)
    qmask = _dialogues(T, N, g)
    params = {k: v.detach() for k, v in model.state_dict().items()}
    ref_b = tp.sps_cell(params, "marn_cell_f", x_l, x_a, qmask)[:, 0]
    ref_1 = tp.sps_cell(params, "marn_cell_f", x_l[:, :1], x_a[:, :1], qmask[:, :1])[:, 0]
    cell = model.to("cuda").marn_cell_f.eval()
    with torch.no_grad():
        our_b = cell(None, x_l.cuda(), x_a.cuda(), qmask.cuda())[:, 0].cpu()
        our_1 = cell(None, x_l[:, :1].cuda().contiguous(), x_a[:, :1].cuda().contiguous(), qmask[:, :1].cuda().contiguous())[:, 0].cpu()
    assert e_inf(our_b, ref_b) < 2e-5 and e_inf(our_1, ref_1) < 2e-5
    assert (ref_b - ref_1).abs().max() > 1e-3 and (our_b - our_1).abs().max() > 1e-3


def test_train_mode_in_kernel_attention_dropout():
    T, N = 6, 20
    model = sps_seeded_model(9, True, "cuda").train()
    x, qmask, umask, labels = _batch(T, N, 4)
    torch.manual_seed(3)
    a, da = _fwd_bwd(model, x, qmask, umask, labels)
    torch.manual_seed(3)
    b, db = _fwd_bwd(model, x, qmask, umask, labels)
    c, _ = _fwd_bwd(model, x, qmask, umask, labels)
    assert torch.equal(a, b) and torch.equal(da, db) and not torch.equal(a, c)
    assert torch.isfinite(da).all()


def test_shard_above_the_cooperative_capacity_fails_loudly():
    """MODE 0 couples the dialogues of a shard (F3) and needs all its CTAs co-resident: above 8 x #SMs dialogues the call must
    raise with an explanation, never split the shard silently (that would change the function)."""
    T, N = 2, 148 * 8 + 8
    model = sps_seeded_model(3, True, "cuda").eval()
    g = torch.Generator().manual_seed(0)
    x_l, x_a = torch.randn(T, N, 100, generator=g).cuda(), torch.randn(T, N, 100, generator=g).cuda()
    with pytest.raises(RuntimeError, match="cannot be split"):
        model.marn_cell_f(None, x_l, x_a, _dialogues(T, N, g).cuda())


@pytest.mark.parametrize("w", [2, 100, 512])
def test_fused_reverse_seq_matches_reference_semantics(w):
    """lsthm_reverse_seq (one kernel each way) vs the oracle's restatement of MARN1_sps._reverse_seq (lsthm_sps.py:396-410),
    values and gradient (the map is its own adjoint), ragged lengths incl. length 1 and full length."""
    L, lens = 11, [11, 1, 4, 7, 11, 2, 10]
    um = torch.zeros(len(lens), L)
    for b, n in enumerate(lens):
        um[b, :n] = 1
    g = torch.Generator().manual_seed(w)
    X = torch.randn(L, len(lens), w, generator=g)
    Xr = X.clone().requires_grad_(True)
    ref = tp.reverse_seq(Xr, um)
    gout = torch.randn(ref.shape, generator=g)
    (ref * gout).sum().backward()
    Xc = X.cuda().requires_grad_(True)
    out = lsthm_b200.lsthm_sps.reverse_seq(Xc, um.cuda())
    (out * gout.cuda()).sum().backward()
    assert torch.equal(out.detach().cpu(), ref.detach()) and torch.equal(Xc.grad.cpu(), Xr.grad)
