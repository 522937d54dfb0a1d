"""GPU suite for the GRU speaker-state cell (lsthm_onlysp — the default model of the reference's train.py — and
lsthm_nsps), all through the C ABI (lsthm_gsp_*):

  1. MARN_cell level: kernel (fwd + BPTT) vs the oracle's torch restatement + autograd, both party-update rules,
     random two-speaker dialogues incl. padded (all-zero qmask) steps, ragged tiles, several tile heights, eval and full
     mask tape (speaker / hidden-state dropout + attention dropout), multi-CTA shards at the production tile height;
  2. MARN1_onlysp / MARN1_nsps modules vs the reference-generated fixtures (eval, perturbed ones-parameters, train tape);
  3. properties at N=1024, T=110 (determinism, tile-height invariance, inference == training forward) and a shard of
     2048 dialogues (no cooperative-launch cap on this family);
  4. dialogues are independent: a dialogue alone equals the same dialogue inside a batch, bit for bit.
"""
import pytest
import torch

from helpers import (TOL_GRAD, TOL_OUT, check_against_fp64_truth, check_against_golden, e_inf, golden_files, gsp_cell_masks,
                     load_golden, sps_run_module, sps_seeded_model)
from oracle import torch_port as tp

pytestmark = pytest.mark.gpu


def _dialogues(T, N, g, pad_from=None):
    q = torch.zeros(T, N, 2)
    s = torch.randint(0, 2, (N,), generator=g)
    for t in range(T):
        flip = torch.rand(N, generator=g) < 0.6
        s = torch.where(flip, 1 - s, s)
        q[t, torch.arange(N), s] = 1
    if pad_from is not None:              # dialogue 0 is shorter: all-zero speaker rows on its padded steps
        q[pad_from:, 0] = 0
    return q


CELL_CASES = [(6, 5, 4, False), (5, 11, 4, True), (7, 9, 8, False), (4, 13, 7, True), (6, 64, 7, False), (5, 64, 7, True),
              (3, 3, 1, False), (5, 6, 2, True), (4, 10, 3, False), (3, 16, 5, False), (4, 12, 6, True)]


@pytest.mark.parametrize("kind", ["onlysp", "nsps"])
@pytest.mark.parametrize("T,N,rows,masked", CELL_CASES)
def test_cell_vs_oracle(kind, T, N, rows, masked):
    g = torch.Generator().manual_seed(T * 100 + N)
    model = sps_seeded_model(300 + N, True, kind=kind)
    cell, pre = model.marn_cell_f, "marn_cell_f"
    x_l, x_a = torch.randn(T, N, 100, generator=g), torch.randn(T, N, 100, generator=g)
    u = torch.randn(T, N, 200, generator=g)
    qmask = _dialogues(T, N, g, pad_from=T - 2 if T > 3 else None)
    dout = torch.randn(T, N, 512, generator=g)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    xl_r, xa_r, u_r = (t.clone().requires_grad_(True) for t in (x_l, x_a, u))
    tape = tp.DropoutTape(7) if masked else None
    if kind == "onlysp":
        ref = tp.onlysp_cell(params, pre, xl_r, xa_r, qmask, tape)      # its GRU input is cat[x_l, x_a]
    else:
        ref = tp.nsps_cell(params, pre, u_r, xl_r, xa_r, qmask, tape)
    (ref * dout).sum().backward()
    model = model.to("cuda").eval()          # dropout comes only from the explicit mask tape below
    cell.rows_per_cta = rows
    if masked:
        cell.mask_override = gsp_cell_masks(tape, pre, T, N, "cuda")
    xl_c, xa_c, u_c = (t.cuda().requires_grad_(True) for t in (x_l, x_a, u))
    out = cell(torch.cat([xl_c, xa_c], -1) if kind == "onlysp" else u_c, xl_c, xa_c, qmask.cuda())
    (out * dout.cuda()).sum().backward()
    errs = {"out": e_inf(out.detach().cpu(), ref.detach()), "dx_l": e_inf(xl_c.grad.cpu(), xl_r.grad),
            "dx_a": e_inf(xa_c.grad.cpu(), xa_r.grad)}
    if kind == "nsps":
        errs["du"] = e_inf(u_c.grad.cpu(), u_r.grad)
    for n, p in model.named_parameters():
        if n.startswith(pre + ".") and params[n].grad is not None:
            assert p.grad is not None, n
            errs[n] = e_inf(p.grad.cpu(), params[n].grad)
        elif n.startswith(pre + "."):
            assert p.grad is None, n          # lstm_q0/q1/lstm_s/gru_l/crossatt_a2l/Wv: never used
    assert errs["out"] <= 2e-5, errs
    assert max(errs.values()) <= 2e-4, sorted(errs.items(), key=lambda kv: -kv[1])[:6]


@pytest.mark.parametrize("path", golden_files("onlysp_*.npz") + golden_files("nsps_*.npz") + golden_files("no_en_*.npz"),
                         ids=lambda p: p.split("/")[-1][:-4])
def test_module_matches_reference_fixture(path):
    fix = load_golden(path)
    logp, loss, dx, grads = sps_run_module(fix)
    # same protocol as lsthm_sps: err(ours, fp64 truth) <= max(1e-4 | 1e-3, 3 x err(reference fp32, fp64 truth))
    errs = check_against_fp64_truth(fix, logp, loss, dx, grads, tol_out=TOL_OUT, tol_grad=TOL_GRAD)
    if not int(fix["perturb"]):
        check_against_golden(fix, logp, loss, dx, grads, tol_out=TOL_OUT, tol_grad=TOL_GRAD)
    print(path.split("/")[-1], errs)


def _batch(T, N, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, N, 1124, generator=g)
    return x.cuda(), _dialogues(T, N, g).cuda(), torch.ones(N, T).cuda(), torch.randint(0, 6, (N * T,), generator=g).cuda()


def _fwd_bwd(model, x, qmask, umask, labels):
    model.zero_grad(set_to_none=True)
    xx = x.clone().requires_grad_(True)
    logp, _, _ = model(xx, qmask, umask)
    tp.masked_loss(logp, labels, umask, "ce").backward()
    return logp.detach(), xx.grad.detach()


@pytest.mark.parametrize("kind", ["onlysp", "nsps"])
def test_properties_at_benchmark_size(kind):
    T, N = 110, 1024
    model = sps_seeded_model(111, True, "cuda", kind=kind).eval()
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


@pytest.mark.parametrize("kind", ["onlysp", "nsps"])
def test_dialogues_are_independent_and_large_shards_run(kind):
    """A dialogue alone == the same dialogue inside a batch (bitwise at the cell level), also for a shard of 2048 dialogues
    (the lsthm_sps cell is capped at 148 x 8 by its cooperative launch; this family is not)."""
    T, N = 6, 2048
    model = sps_seeded_model(5, True, "cuda", kind=kind).eval()
    cell = model.marn_cell_f
    g = torch.Generator().manual_seed(1)
    x_l, x_a, u = (torch.randn(T, N, d, generator=g).cuda() for d in (100, 100, 200))
    qmask = _dialogues(T, N, g).cuda()
    with torch.no_grad():
        full = cell(u, x_l, x_a, qmask)
        for d in (0, 1029, 2047):
            one = cell(u[:, d:d + 1].contiguous(), x_l[:, d:d + 1].contiguous(), x_a[:, d:d + 1].contiguous(), qmask[:, d:d + 1].contiguous())
            assert torch.equal(one[:, 0], full[:, d]), d
    p = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    sl = slice(1000, 1003)
    ref = tp.nsps_cell(p, "marn_cell_f", u[:, sl].cpu(), x_l[:, sl].cpu(), x_a[:, sl].cpu(), qmask[:, sl].cpu()) if kind == "nsps" \
        else None
    if ref is not None:
        assert e_inf(full[:, sl].cpu(), ref) < 2e-5


@pytest.mark.parametrize("kind", ["onlysp", "nsps"])
def test_train_mode_in_kernel_attention_dropout(kind):
    T, N = 6, 20
    model = sps_seeded_model(9, True, "cuda", kind=kind).train()
    x, qmask, umask, labels = _batch(T, N, 4)
    torch.manual_seed(3)
    a, da = _fwd_bwd(model, x, qmask, umask, labels)
    torch.manual_seed(3)
    b, db = _fwd_bwd(model, x, qmask, umask, labels)
    c, _ = _fwd_bwd(model, x, qmask, umask, labels)
    assert torch.equal(a, b) and torch.equal(da, db) and not torch.equal(a, c)
    assert torch.isfinite(da).all()


@pytest.mark.parametrize("kind", ["onlysp", "sps"])
def test_in_kernel_attention_dropout_is_the_same_mask_forward_and_backward(kind):
    """Production train mode draws the in-cell attention dropout from a counter hash inside the kernels (no mask tensor):
    forward and backward must evaluate the same mask.  With the seed fixed the cell is a deterministic function, so the
    directional derivative of sum(out * w) along a random direction must equal <grad, direction>; a mismatch between the two
    kernels' masks would show up as an O(1) error."""
    from importlib import import_module
    import lsthm_b200
    T, N = 3, 6
    model = sps_seeded_model(21, True, "cuda", kind=kind).eval()       # recurrent-state dropouts off; attention dropout via opts below
    cell = model.marn_cell_f
    g = torch.Generator().manual_seed(2)
    gx = (0.5 * torch.randn(T, N, 2, 512, generator=g)).cuda()
    qmask = _dialogues(T, N, g).cuda()
    w = torch.randn(T, N, 512, generator=g).cuda()
    dirn = torch.randn(gx.shape, generator=g).cuda()
    if kind == "onlysp":
        rec = import_module(lsthm_b200.__name__ + ".gsp_recurrence")
        gxs = (0.5 * torch.randn(T, N, 384, generator=g)).cuda()
        f = lambda a: rec.gsp_cell(a, gxs, qmask, (None,) * 4, cell.cell_weights(), 0, 0, 0.2, 12345)
    else:
        rec = import_module(lsthm_b200.__name__ + ".sps_recurrence")
        f = lambda a: rec.sps_cell(a, qmask, (None,) * 5, cell.cell_weights(), 0, 0.2, 12345)
    a0 = gx.clone().requires_grad_(True)
    out = f(a0)
    (out * w).sum().backward()
    assert torch.equal(out.detach(), f(gx).detach())                                  # seeded
    assert not torch.equal(out.detach()[..., 256:384], model.eval() and
                           (rec.gsp_cell(gx, gxs, qmask, (None,) * 4, cell.cell_weights(), 0, 0, 0.0, 0) if kind == "onlysp"
                            else rec.sps_cell(gx, qmask, (None,) * 5, cell.cell_weights(), 0, 0.0, 0))[..., 256:384])   # dropout is active
    eps = 2e-3
    with torch.no_grad():
        num = ((f(gx + eps * dirn) * w).double().sum() - (f(gx - eps * dirn) * w).double().sum()) / (2 * eps)
    ana = (a0.grad * dirn).double().sum()
    assert abs(float(num - ana)) <= 5e-3 * max(abs(float(ana)), 1.0), (float(num), float(ana))


@pytest.mark.parametrize("kind", ["onlysp", "nsps", "sps"])
def test_two_stream_encoders_change_nothing_but_the_schedule(kind):
    """The text and audio encoder chains run on two CUDA streams, forward and backward (streams.fork_join).  Every kernel is
    deterministic, so log-probs, dx and every parameter gradient must be BITWISE those of the single-stream schedule — a missing
    fork/join dependency or memory reused too early shows up as a difference.  Repeated, at a size where the chains overlap."""
    T, N = 40, 96
    model = sps_seeded_model(31, True, "cuda", kind=kind).eval()
    assert model.concurrent_encoders
    x, qmask, umask, labels = _batch(T, N, 5)

    def run():
        logp, dx = _fwd_bwd(model, x, qmask, umask, labels)
        return logp, dx, {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    runs = [run() for _ in range(3)]
    model.concurrent_encoders = False
    l0, dx0, g0 = run()
    for l1, dx1, g1 in runs:
        assert torch.equal(l0, l1) and torch.equal(dx0, dx1)
        bad = [k for k in g0 if not torch.equal(g0[k], g1[k])]
        assert g0.keys() == g1.keys() and not bad, bad


@pytest.mark.parametrize("kind", ["onlysp", "no_en", "sps"])
def test_no_grad_inference_takes_the_same_kernels(kind):
    T, N = 9, 6
    model = sps_seeded_model(32, True, "cuda", kind=kind).eval()
    x, qmask, umask, _ = _batch(T, N, 6)
    lp_grad = model(x, qmask, umask)[0]
    with torch.no_grad():
        lp0 = model(x, qmask, umask)[0]
    assert lp_grad.requires_grad and not lp0.requires_grad and torch.equal(lp0, lp_grad.detach())
