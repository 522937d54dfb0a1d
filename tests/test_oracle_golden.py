"""CPU suite, part 1: pin the oracle.

* torch restatement (oracle/torch_port.py) vs the committed reference-generated fixtures
  (tests/golden/, made by oracle/make_golden.py from the live reference);
* the same against the live reference when /root/reference is present (build container only);
* plain-C restatement (oracle/mab_oracle.c, fp64) of the recurrence forward and of the hand-derived
  BPTT vs the torch restatement + autograd.
"""
import numpy as np
import pytest
import torch

from helpers import SPEC, check_against_golden, golden_files, load_golden, port_run, seeded_model
from oracle import cpu as ocpu
from oracle import torch_port as tp
from oracle.ref_shim import attach_tape, load_reference, reference_available


@pytest.mark.parametrize("path", golden_files("sps_*.npz") + golden_files("onlysp_*.npz") + golden_files("nsps_*.npz")
                         + golden_files("no_en_*.npz"),
                         ids=lambda p: p.split("/")[-1][:-4])
def test_sps_torch_port_matches_reference_fixture(path):
    """lsthm_sps: the index-form restatement (vectorised _select_parties, no per-row loop) vs the reference
    fixtures, eval / perturbed ones-parameters / train with the full dropout mask tape."""
    from helpers import sps_port_run
    fix = load_golden(path)
    logp, loss, dx, grads = sps_port_run(fix)
    check_against_golden(fix, logp, loss, dx, grads, tol_out=1e-5, tol_grad=2e-4)


def test_sps_plan_and_reverse_match_reference_semantics():
    from importlib import import_module
    import lsthm_b200
    sr = import_module(lsthm_b200.__name__ + ".sps_recurrence")
    g = torch.Generator().manual_seed(0)
    q = torch.zeros(6, 9, 2)
    r = torch.rand(6, 9, generator=g)
    q[..., 0] = (r < 0.45).float()
    q[..., 1] = ((r >= 0.45) & (r < 0.9)).float()          # the rest are padded rows [0,0] -> speaker 0
    q[2, :, 0], q[2, :, 1] = 1, 0                            # a step where nobody is speaker 1
    pi, pr, n0 = sr.party_plan(q)
    for t in range(6):
        P0, P1 = tp.party_rows(q[t])
        assert torch.equal(pi[t], torch.cat([P0, P1]).int()) and int(n0[t]) == P0.numel()
        assert torch.equal(pr[t][pi[t].long()], torch.arange(9).int())
    um = torch.zeros(4, 6)
    for b, n in enumerate([6, 2, 4, 1]):
        um[b, :n] = 1
    X = torch.randn(6, 4, 5, generator=g)
    assert torch.equal(lsthm_b200.lsthm_sps.reverse_seq(X, um), tp.reverse_seq(X, um))


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_torch_port_matches_reference_fixture(path):
    fix = load_golden(path)
    probs, loss, dx, grads = port_run(fix)
    errs = check_against_golden(fix, probs, loss, dx, grads, tol_out=2e-6, tol_grad=2e-5)
    assert errs["probs"] < 2e-6


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("kind", ["ATV", "AT"])
def test_torch_port_matches_live_reference(kind):
    ref = load_reference()
    cls = ref.MARN_ATV if kind == "ATV" else ref.MARN_AT
    torch.manual_seed(7)
    m = cls()
    ours = seeded_model(kind, 7)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), ours.state_dict().values()))
    x = torch.randn(8, 4, SPEC[kind]["din"])
    tape = tp.DropoutTape(3)
    attach_tape(m, tape)
    m.train()
    y = m(x)
    y2 = tp.mab_forward({k: v.detach() for k, v in m.state_dict().items()}, x, kind, tape.rewind())
    assert (y - y2).abs().max() < 1e-6


@pytest.mark.parametrize("kind", ["ATV", "AT"])
def test_c_oracle_matches_torch_port_and_autograd(kind):
    torch.set_default_dtype(torch.float64)
    try:
        model = seeded_model(kind, 5).double()
        p = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
        T, N = 6, 3
        dh, rd = SPEC[kind]["dh"], SPEC[kind]["rd"]
        x = torch.randn(T, N, SPEC[kind]["din"])
        tape = tp.DropoutTape(1)
        _, hz0, xs = tp.mab_forward(p, x, kind, tape, return_state=True)
        gx = tp.mab_gate_inputs(p, xs, kind).detach().requires_grad_(True)
        hz = tp.mab_recurrence(p, gx, kind, tape.rewind())
        assert (hz - hz0).abs().max() < 1e-12          # hoisted W.x == per-step form
        mask = tape.stacked("fc.2").numpy()
        pn = {k: v.detach().numpy() for k, v in p.items()}
        f = ocpu.mab_forward(pn, gx.detach().numpy(), dh, rd, mask)
        assert np.abs(f["hz"] - hz.detach().numpy()).max() < 1e-12
        g = torch.randn_like(hz)
        names = [k for k in p if k.split(".")[0] in ("att", "fc") or k.startswith("reduce")
                 or k.endswith(".U.weight") or k.endswith(".V.weight")]
        gr = torch.autograd.grad((hz * g).sum(), [gx] + [p[k] for k in names])
        adj, grads = ocpu.mab_backward(pn, g.numpy(), f, dh, rd, mask)
        assert np.abs(adj["dgx"] - gr[0].numpy()).max() < 1e-12
        for k, gg in zip(names, gr[1:]):
            assert np.abs(grads[k] - gg.numpy()).max() < 1e-11, k
    finally:
        torch.set_default_dtype(torch.float32)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("variant", ["onlysp", "nsps", "no_en"])
@pytest.mark.parametrize("train,perturb", [(False, False), (False, True), (True, True)])
def test_gru_variant_port_matches_live_reference(variant, train, perturb):
    """lsthm_onlysp (the reference's train.py default model) and lsthm_nsps (listener update, softmax(p) fusion): the
    oracle restatement vs the live reference — log-probs, loss and every gradient, eval / perturbed ones-parameters /
    train with the mask tape, ragged dialogue lengths."""
    from oracle.make_golden import synth_dialogues
    ref = load_reference()
    torch.manual_seed(31)
    m = {"onlysp": lambda: ref.MARN1_onlysp(6), "nsps": lambda: ref.MARN1_nsps(6, "IEMOCAP"),
         "no_en": lambda: ref.MARN1_no_en(6, "IEMOCAP")}[variant]()
    forward = {"onlysp": tp.onlysp_forward, "nsps": tp.nsps_forward, "no_en": tp.no_en_forward}[variant]
    if perturb:
        tp.perturb_ones(m, 5)
    x, qmask, umask, labels = synth_dialogues(17, 9, [9, 4, 7, 9, 5])
    x.requires_grad_(True)
    tape = None
    if train:
        tape = tp.DropoutTape(4)
        attach_tape(m, tape)
        m.train()
    else:
        m.eval()
    logp, _, _ = m(x, qmask, umask)
    loss = ref.MaskedLoss(torch.nn.CrossEntropyLoss)(logp, labels.view(-1), umask)
    loss.backward()
    p = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    x2 = x.detach().clone().requires_grad_(True)
    logp2, _, _ = forward(p, x2, qmask, umask, tape.rewind() if train else None)
    loss2 = tp.masked_loss(logp2, labels.view(-1), umask, "ce")
    loss2.backward()
    e = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert e(logp2.detach(), logp.detach()) < 2e-5 and abs(float(loss2.detach()) - float(loss.detach())) / abs(float(loss.detach())) < 1e-5
    assert e(x2.grad, x.grad) < 5e-4
    gmax = max(float(q.grad.norm()) for q in m.parameters() if q.grad is not None)
    for n, q in m.named_parameters():
        if q.grad is None:
            assert p[n].grad is None or float(p[n].grad.abs().max()) == 0.0, n
        elif float(q.grad.norm()) > 1e-6 * gmax:
            assert e(p[n].grad, q.grad) < 1e-3, (n, e(p[n].grad, q.grad))


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("variant", ["sps", "onlysp", "nsps", "no_en"])
def test_speaker_family_state_dict_and_init_match_live_reference(variant):
    """Drop-in boundary: parameter names, shapes, registration order — including the never-used tensors — and
    the default-init RNG order are those of the reference, so checkpoints load both ways and seed_everything gives equal weights."""
    import lsthm_b200
    ref = load_reference()
    ours = {"sps": lambda: lsthm_b200.lsthm_sps.MARN1_sps(6), "onlysp": lambda: lsthm_b200.lsthm_onlysp.MARN1_onlysp(6),
            "nsps": lambda: lsthm_b200.lsthm_nsps.MARN1_nsps(6, "IEMOCAP"),
            "no_en": lambda: lsthm_b200.lsthm_no_en.MARN1_no_en(6, "IEMOCAP")}[variant]
    theirs = {"sps": lambda: ref.MARN1_sps(6), "onlysp": lambda: ref.MARN1_onlysp(6), "nsps": lambda: ref.MARN1_nsps(6, "IEMOCAP"),
              "no_en": lambda: ref.MARN1_no_en(6, "IEMOCAP")}[variant]
    torch.manual_seed(111)
    a = ours()
    torch.manual_seed(111)
    b = theirs()
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]) for k in sa)
    b.load_state_dict(sa)
    a.load_state_dict(sb)
