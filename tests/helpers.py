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
    for key in fix:
        if key.startswith("gnone/"):
            assert grads.get(key[6:]) is None, f"{key[6:]} must have no gradient (SURVEY.md F8)"
        if not key.startswith("gsamp/"):
            continue
        name = key[6:]
        g = grads[name]
        assert g is not None, name
        flat = np.asarray(g.detach().cpu()).reshape(-1)
        samp = flat if flat.size <= 4096 else flat[::stride]
        scale = float(fix["gnorm/" + name]) / np.sqrt(flat.size) + 1e-30   # rms of the reference grad
        err = float(np.abs(samp - fix[key]).max() / max(np.abs(fix[key]).max(), scale))
        nerr = abs(float(np.linalg.norm(flat.astype(np.float64))) - float(fix["gnorm/" + name])) / (float(fix["gnorm/" + name]) + 1e-30)
        worst = max(worst, err, nerr)
        assert err <= tol_grad and nerr <= tol_grad, (name, err, nerr)
    errs["grads"] = worst
    return errs


def attach_tape_to_ours(model, tape):
    """Encoder dropouts are ordinary nn.Dropout modules in our mirror: drive them from the tape."""
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
            if isinstance(child, nn.Dropout) and path.startswith("encoder"):
                setattr(mod, cname, _TapeDropout(f"{path}.{cname}", child.p))


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
