"""Shared test helpers: golden fixtures, error metrics (scale-relative, never
element-wise relative), oracle drivers."""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

import lsthm_b200  # noqa: E402
from oracle import torch_port as tp  # noqa: E402

SPEC = {"ATV": dict(cls=lambda: lsthm_b200.HybridRNN_ATV.MARN, din=712, C=6, dh=(128, 16, 64), rd=(16, 128, 100)),
        "AT": dict(cls=lambda: lsthm_b200.HybridRNN_AT.MARN, din=200, C=7, dh=(128, 16), rd=(16, 128))}

# fp32 parity bars of BASELINE.json:north_star — logits/loss 1e-4, gradients 1e-3 (scale-relative)
TOL_OUT, TOL_GRAD = 1e-4, 1e-3


def e_inf(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def e_2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def golden_files(pattern="mab_*.npz"):
    return sorted(glob.glob(os.path.join(GOLDEN, pattern)))


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def seeded_model(kind, seed, device="cpu"):
    """Our drop-in module with the default init under the fixture's seed == the reference's weights."""
    torch.manual_seed(int(seed))
    return SPEC[kind]["cls"]()().to(device)


def tape_from_fixture(fix):
    tape = tp.DropoutTape(0)
    for k, v in fix.items():
        if k.startswith("tape/"):
            tape.masks[k[5:]] = [torch.from_numpy(m) for m in v]
    for k in sorted(k for k in fix if k.startswith("tapei/")):      # per-call masks of a site whose calls differ in shape
        tape.masks.setdefault(k[6:].rsplit("/", 1)[0], []).append(torch.from_numpy(fix[k]))
    return tape.rewind()


def masked_ce(probs, labels, T, N):
    return tp.masked_loss(probs, labels, torch.ones(N, T, device=probs.device), "ce")


def port_run(fix, params=None):
    """Run the oracle's torch restatement on a fixture; returns probs, loss, dx, {name: grad}."""
    kind = str(fix["kind"])
    if params is None:
        model = seeded_model(kind, fix["seed"])
        params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    params = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    x = torch.from_numpy(fix["x"]).clone().requires_grad_(True)
    tape = tape_from_fixture(fix) if int(fix["train"]) else None
    probs = tp.mab_forward(params, x, kind, tape)
    loss = masked_ce(probs, torch.from_numpy(fix["labels"]), int(fix["T"]), int(fix["N"]))
    loss.backward()
    return probs.detach(), loss.detach(), x.grad, {k: v.grad for k, v in params.items()}


def check_against_golden(fix, probs, loss, dx, grads, tol_out=TOL_OUT, tol_grad=TOL_GRAD):
    """Assert the parity bar against a reference-generated fixture.  grads: name -> tensor or None."""
    errs = {"probs": e_inf(probs, fix["probs"]), "loss": abs(float(loss) - float(fix["loss"])) / abs(float(fix["loss"])),
            "dx": e_inf(dx, fix["dx"])}
    assert errs["probs"] <= tol_out, errs
    assert errs["loss"] <= tol_out, errs
    assert errs["dx"] <= tol_grad, errs
    assert (np.argmax(np.asarray(probs), -1) == np.argmax(fix["probs"], -1)).all()
    stride = int(fix["sample_stride"])
    worst = 0.0
    gmax = max(float(fix[k]) for k in fix if k.startswith("gnorm/"))
    for key in fix:
        if key.startswith("gnone/"):
            assert grads.get(key[6:]) is None, f"{key[6:]} must have no gradient (SURVEY.md F8)"
        if not key.startswith("gsamp/"):
            continue
        name = key[6:]
        g = grads[name]
        assert g is not None, name
        flat = np.asarray(g.detach().cpu()).reshape(-1)
        if float(fix["gnorm/" + name]) < 1e-6 * gmax:
            # analytically-zero gradient (e.g. in-cell Wq/Wk while Wk is all ones: the softmax is uniform):
            # the reference holds rounding noise there; require ours to be noise-sized too
            assert float(np.linalg.norm(flat.astype(np.float64))) < 1e-5 * gmax, name
            continue
        samp = flat if flat.size <= 4096 else flat[::stride]
        scale = float(fix["gnorm/" + name]) / np.sqrt(flat.size) + 1e-30   # rms of the reference grad
        err = float(np.abs(samp - fix[key]).max() / max(np.abs(fix[key]).max(), scale))
        nerr = abs(float(np.linalg.norm(flat.astype(np.float64))) - float(fix["gnorm/" + name])) / (float(fix["gnorm/" + name]) + 1e-30)
        worst = max(worst, err, nerr)
        assert err <= tol_grad and nerr <= tol_grad, (name, err, nerr)
    errs["grads"] = worst
    return errs


def attach_tape_to_ours(model, tape, skip_prefixes=None):
    """Dropouts outside the fused kernels are ordinary nn.Dropout modules in our mirror: drive them from
    the tape (site name = module path, as oracle/ref_shim.attach_tape does for the reference)."""
    import torch.nn as nn

    class _TapeDropout(nn.Module):
        def __init__(self, site, p):
            super().__init__()
            self.site, self.p = site, p

        def forward(self, x):
            if not self.training or self.p == 0.0:
                return x
            return x * tape.mask(self.site, x.shape, self.p, x.dtype).to(x.device)

    for path, mod in list(model.named_modules()):
        for cname, child in list(mod.named_children()):
            site = f"{path}.{cname}" if path else cname
            if not isinstance(child, nn.Dropout):
                continue
            if skip_prefixes is None:
                if path.startswith("encoder"):
                    setattr(mod, cname, _TapeDropout(site, child.p))
            elif not site.startswith(tuple(skip_prefixes)):
                setattr(mod, cname, _TapeDropout(site, child.p))


def run_module(fix, device="cpu", rows_per_cta=0):
    """Run OUR drop-in module on a fixture (weights from the fixture's seed); fwd + loss + bwd."""
    kind = str(fix["kind"])
    model = seeded_model(kind, fix["seed"], device)
    model.rows_per_cta = rows_per_cta
    T, N = int(fix["T"]), int(fix["N"])
    x = torch.from_numpy(fix["x"]).to(device).requires_grad_(True)
    if int(fix["train"]):
        model.train()
        tape = tape_from_fixture(fix)
        attach_tape_to_ours(model, tape)
        model.fc_mask_override = tape.stacked("fc.2").to(device)
    else:
        model.eval()
    probs = model(x)
    loss = masked_ce(probs, torch.from_numpy(fix["labels"]).to(device), T, N)
    loss.backward()
    grads = {n: (None if p.grad is None else p.grad.detach().cpu()) for n, p in model.named_parameters()}
    return probs.detach().cpu(), loss.detach().cpu(), x.grad.detach().cpu(), grads


# ------------------------------------------------------------------------------------------------
# lsthm_sps helpers
# ------------------------------------------------------------------------------------------------
SPEAKER_MODELS = {"sps": lambda: lsthm_b200.lsthm_sps.MARN1_sps(6), "onlysp": lambda: lsthm_b200.lsthm_onlysp.MARN1_onlysp(6),
                  "nsps": lambda: lsthm_b200.lsthm_nsps.MARN1_nsps(6, "IEMOCAP"),
                  "no_en": lambda: lsthm_b200.lsthm_no_en.MARN1_no_en(6, "IEMOCAP")}
SPEAKER_PORTS = {"sps": tp.sps_forward, "onlysp": tp.onlysp_forward, "nsps": tp.nsps_forward, "no_en": tp.no_en_forward}


def sps_seeded_model(seed, perturb, device="cpu", kind="sps"):
    torch.manual_seed(int(seed))
    m = SPEAKER_MODELS[kind]()
    if perturb:
        tp.perturb_ones(m, int(seed) + 3)
    return m.to(device)


def sps_cell_masks(tape, pre, qmask_dir, device="cpu"):
    """Turn the tape's sequential calls of ``<pre>.dropout`` (order per step: hq0 if N0, hq1 if N1, h_l, h_a;
    model/lsthm_sps.py:184,189,211,213) and ``<pre>.crossatt_l2a.dropout`` into the [T,N,...] mask tensors the
    kernel consumes.  qmask_dir is the qmask that cell saw (reversed for the backward cell)."""
    T, N, _ = qmask_dir.shape
    calls = tape.masks[pre + ".dropout"]
    mq0, mq1, ml, ma = (torch.ones(T, N, 128) for _ in range(4))
    i = 0
    for t in range(T):
        P0, P1 = tp.party_rows(qmask_dir[t])
        if P0.numel():
            mq0[t] = calls[i]; i += 1
        if P1.numel():
            mq1[t] = calls[i]; i += 1
        ml[t] = calls[i]; ma[t] = calls[i + 1]; i += 2
    assert i == len(calls)
    att = torch.stack([m.float() for m in tape.masks[pre + ".crossatt_l2a.dropout"]], 0)
    return tuple(x.to(device).contiguous() for x in (mq0, mq1, ml, ma, att))


def gsp_cell_masks(tape, pre, T, N, device="cpu"):
    """GRU-variant cells (lsthm_onlysp.py:176,184,186 / lsthm_nsps.py:183,192,194): ``<pre>.dropout`` is called three times per
    step — on h_s, h_l, h_a — and ``<pre>.crossatt_l2a.dropout`` once: -> the (ms, ml, ma, att_mask) tensors of the kernel."""
    calls = tape.masks[pre + ".dropout"]
    assert len(calls) == 3 * T
    ms, ml, ma = (torch.stack([calls[3 * t + i] for t in range(T)], 0).float() for i in range(3))
    att = torch.stack([m.float() for m in tape.masks[pre + ".crossatt_l2a.dropout"]], 0)
    return tuple(x.to(device).contiguous() for x in (ms, ml, ma, att))


def sps_port_run(fix):
    kind = str(fix["kind"]) if "kind" in fix else "sps"
    model = sps_seeded_model(fix["seed"], int(fix["perturb"]), kind=kind)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    x = torch.from_numpy(fix["x"]).clone().requires_grad_(True)
    qmask, umask = torch.from_numpy(fix["qmask"]), torch.from_numpy(fix["umask"])
    tape = tape_from_fixture(fix) if int(fix["train"]) else None
    logp, x_l, x_a = SPEAKER_PORTS[kind](params, x, qmask, umask, tape)
    loss = tp.masked_loss(logp, torch.from_numpy(fix["labels"]).view(-1), umask, "ce")
    loss.backward()
    return logp.detach(), loss.detach(), x.grad, {k: v.grad for k, v in params.items()}


def sps_run_module(fix, device="cuda", rows_per_cta=0):
    kind = str(fix["kind"]) if "kind" in fix else "sps"
    model = sps_seeded_model(fix["seed"], int(fix["perturb"]), device, kind=kind)
    model.marn_cell_f.rows_per_cta = model.marn_cell_b.rows_per_cta = rows_per_cta
    x = torch.from_numpy(fix["x"]).to(device).requires_grad_(True)
    qmask, umask = torch.from_numpy(fix["qmask"]).to(device), torch.from_numpy(fix["umask"]).to(device)
    if int(fix["train"]):
        model.train()
        tape = tape_from_fixture(fix)
        q_cpu, u_cpu = torch.from_numpy(fix["qmask"]), torch.from_numpy(fix["umask"])
        if kind == "sps":
            model.marn_cell_f.mask_override = sps_cell_masks(tape, "marn_cell_f", q_cpu, device)
            model.marn_cell_b.mask_override = sps_cell_masks(tape, "marn_cell_b", tp.reverse_seq(q_cpu, u_cpu), device)
        else:
            T_, N_ = q_cpu.shape[0], q_cpu.shape[1]
            model.marn_cell_f.mask_override = gsp_cell_masks(tape, "marn_cell_f", T_, N_, device)
            model.marn_cell_b.mask_override = gsp_cell_masks(tape, "marn_cell_b", T_, N_, device)
        attach_tape_to_ours(model, tape, skip_prefixes=("marn_cell",))
    else:
        model.eval()
    logp, _, _ = model(x, qmask, umask)
    loss = tp.masked_loss(logp, torch.from_numpy(fix["labels"]).view(-1).to(device), umask, "ce")
    loss.backward()
    grads = {n: (None if p.grad is None else p.grad.detach().cpu()) for n, p in model.named_parameters()}
    return logp.detach().cpu(), loss.detach().cpu(), x.grad.detach().cpu(), grads


class relu_kinks:
    """Records every ``torch.relu`` pre-activation of an oracle run (call order = site id) and can force the on/off
    decision of chosen units: ``force = {site: [(flat_index, on), ...]}``.  A ReLU whose pre-activation is closer to
    zero than the forward parity tolerance has no defined derivative at that resolution; the sps parity test uses
    this to enumerate the subgradient choices the fp64 truth could equally have made (see ``sps_kink_truths``)."""

    def __init__(self, force=None):
        self.force, self.z = force or {}, []

    def __enter__(self):
        self._orig = torch.relu

        def patched(z):
            i = len(self.z)
            self.z.append(z.detach())
            m = z > 0
            if i in self.force:
                m = m.clone()
                for idx, on in self.force[i]:
                    m.view(-1)[idx] = on
            return z * m.to(z.dtype)
        torch.relu = patched
        return self

    def __exit__(self, *exc):
        torch.relu = self._orig


def sps_port_run64(fix, force=None):
    """fp64 run of the oracle restatement (eval fixtures only) -> truth dict shaped like the fixture's fp64 fields,
    plus the recorded ReLU pre-activations."""
    assert not int(fix["train"])
    kind = str(fix["kind"]) if "kind" in fix else "sps"
    model = sps_seeded_model(fix["seed"], int(fix["perturb"]), kind=kind)
    params = {k: v.detach().double().requires_grad_(True) for k, v in model.state_dict().items()}
    x = torch.from_numpy(fix["x"]).double().requires_grad_(True)
    qmask, umask = torch.from_numpy(fix["qmask"]).double(), torch.from_numpy(fix["umask"]).double()
    with relu_kinks(force) as rk:
        logp, _, _ = SPEAKER_PORTS[kind](params, x, qmask, umask, None)
        loss = tp.masked_loss(logp, torch.from_numpy(fix["labels"]).view(-1), umask, "ce")
        loss.backward()
    stride = int(fix["sample_stride"])
    truth = {"probs64": logp.detach().numpy(), "loss64": np.array(loss.item()), "dx64": x.grad.numpy()}
    for k, v in params.items():
        if v.grad is None:
            continue
        flat = v.grad.reshape(-1).numpy()
        truth["gnorm64/" + k] = np.array(np.linalg.norm(flat))
        truth["gsamp64/" + k] = flat if flat.size <= 4096 else flat[::stride]
    return truth, rk.z


def sps_kink_truths(fix, rel=TOL_OUT, kmax=5):
    """Alternative fp64 truths of an eval fixture: one per non-empty subset of the (at most kmax) ReLU units whose
    pre-activation lies within ``rel`` x the layer's largest pre-activation of zero — i.e. within the FORWARD parity
    tolerance of their kink — with those units' on/off decisions flipped."""
    import itertools
    base, zs = sps_port_run64(fix)
    assert e_inf(base["dx64"], fix["dx64"]) < 1e-7 and e_inf(base["probs64"], fix["probs64"]) < 1e-9   # oracle == reference (fp64)
    near = []
    for site, z in enumerate(zs):
        a = z.abs().reshape(-1)
        thr = rel * float(a.max())
        for idx in torch.nonzero(a < thr).reshape(-1).tolist():
            near.append((float(a[idx]) / float(a.max()), site, idx, bool(z.reshape(-1)[idx] > 0)))
    near = sorted(near)[:kmax]
    for r in range(1, len(near) + 1):
        for sub in itertools.combinations(near, r):
            force = {}
            for _, site, idx, on in sub:
                force.setdefault(site, []).append((idx, not on))
            yield sub, sps_port_run64(fix, force)[0]


def check_against_fp64_truth(fix, probs, loss, dx, grads, tol_out=TOL_OUT, tol_grad=TOL_GRAD, slack=3.0, truth=None):
    """Parity bar for the ill-conditioned lsthm_sps: against the fp64 run of the reference,
    err(ours, fp64) <= max(tol, slack * err(reference fp32, fp64)), per tensor.  Also requires identical argmax
    with the fp32 reference wherever its top-2 margin exceeds 1e-3.  ``truth`` (optional) replaces the fp64
    comparator (a kink-flipped truth from ``sps_kink_truths``); the bars always come from the fixture."""
    def bar(tol, ref32, truth_):
        return max(tol, slack * e_inf(ref32, truth_))
    bars = {"probs": bar(tol_out, fix["probs"], fix["probs64"]), "dx": bar(tol_grad, fix["dx"], fix["dx64"])}
    gbars = {k[8:]: bar(tol_grad, fix["gsamp/" + k[8:]], fix[k]) for k in fix if k.startswith("gsamp64/")}
    orig = fix
    if truth is not None:
        fix = dict(fix)
        fix.update(truth)
    errs = {"probs": e_inf(probs, fix["probs64"]), "dx": e_inf(dx, fix["dx64"]),
            "loss": abs(float(loss) - float(fix["loss64"])) / abs(float(fix["loss64"]))}
    assert errs["probs"] <= bars["probs"], errs
    assert errs["dx"] <= bars["dx"], errs
    assert errs["loss"] <= max(tol_out, slack * abs(float(orig["loss"]) - float(orig["loss64"])) / abs(float(orig["loss64"]))), errs
    ref = fix["probs"]
    top2 = np.sort(ref, -1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 1e-3
    assert (np.argmax(np.asarray(probs), -1)[decided] == np.argmax(ref, -1)[decided]).all()
    stride = int(fix["sample_stride"])
    gmax = max(float(fix[k]) for k in fix if k.startswith("gnorm64/"))
    worst = 0.0
    for key in fix:
        if key.startswith("gnone64/"):
            assert grads.get(key[8:]) is None, f"{key[8:]} must have no gradient (SURVEY.md F8)"
        if not key.startswith("gsamp64/"):
            continue
        name = key[8:]
        g = grads[name]
        assert g is not None, name
        flat = np.asarray(g.detach().cpu()).reshape(-1)
        if float(fix["gnorm64/" + name]) < 1e-6 * gmax:
            assert float(np.linalg.norm(flat.astype(np.float64))) < 1e-5 * gmax, name
            continue
        samp = flat if flat.size <= 4096 else flat[::stride]
        err = e_inf(samp, fix[key])
        lim = gbars[name]
        worst = max(worst, err / lim)
        assert err <= lim, (name, err, lim)
    errs["worst_grad_over_bar"] = worst
    return errs
