"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the LIVE reference
(/root/reference, imported through oracle/ref_shim.py) on seeded inputs.  Runs only in the
build container (the reference tree does not exist on the GPU box); the fixtures it writes are
committed, and every consumer regenerates the *weights* from the recorded seed (module default
init under torch.manual_seed, verified identical between the reference classes and ours).

    python oracle/make_golden.py            # rewrites all fixtures

What is pinned per case: output probabilities, the MaskedLoss(CrossEntropy) scalar
(loss.py:13-21), d loss/d x, and for every parameter its gradient L2 norm plus a strided sample
(full gradients for tensors <= 4096 elements).  Train-mode cases also store the dropout masks of
every nn.Dropout site (mask tape).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_shim import attach_tape, load_reference  # noqa: E402
from oracle.torch_port import DropoutTape  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SAMPLE_STRIDE = 61
LABEL_P = np.array([144, 245, 384, 170, 299, 381], dtype=np.float64) / 1623.0   # res.csv class frequencies


def grad_summary(named_grads):
    out = {}
    for name, g in named_grads:
        if g is None:
            out["gnone/" + name] = np.zeros(0, np.float32)
            continue
        flat = g.detach().reshape(-1)
        out["gnorm/" + name] = np.array(flat.double().norm().item())
        out["gsamp/" + name] = (flat if flat.numel() <= 4096 else flat[::SAMPLE_STRIDE]).numpy().copy()
    return out


def mab_case(ref, kind, seed, T, N, train):
    cls, din, C = (ref.MARN_ATV, 712, 6) if kind == "ATV" else (ref.MARN_AT, 200, 7)
    torch.manual_seed(seed)
    model = cls()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(T, N, din, generator=g).requires_grad_(True)
    p = LABEL_P if C == 6 else np.full(7, 1 / 7)
    labels = torch.from_numpy(np.random.default_rng(seed).choice(C, size=T * N, p=p)).long()
    mask = torch.ones(N, T)
    fix = dict(kind=kind, seed=seed, T=T, N=N, train=int(train), x=x.detach().numpy().copy(),
               labels=labels.numpy(), sample_stride=SAMPLE_STRIDE)
    if train:
        tape = DropoutTape(seed + 2)
        attach_tape(model, tape)
        model.train()
    else:
        model.eval()
    probs = model(x)
    loss = ref.MaskedLoss(torch.nn.CrossEntropyLoss)(probs, labels, mask)
    loss.backward()
    fix["probs"] = probs.detach().numpy().copy()
    fix["loss"] = np.array(loss.item())
    fix["dx"] = x.grad.numpy().copy()
    fix.update(grad_summary((n, q.grad) for n, q in model.named_parameters()))
    if train:
        for site, masks in tape.masks.items():
            fix["tape/" + site] = torch.stack(masks, 0).numpy().astype(np.float32)
    return fix


def main():
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    cases = [("ATV", 111, 7, 3, False), ("ATV", 112, 5, 2, True), ("AT", 111, 7, 3, False), ("AT", 113, 5, 2, True),
             ("ATV", 114, 12, 9, False)]
    for kind, seed, T, N, train in cases:
        fix = mab_case(ref, kind, seed, T, N, train)
        name = f"mab_{kind}_s{seed}_T{T}_N{N}_{'train' if train else 'eval'}.npz"
        np.savez_compressed(os.path.join(OUT, name), **fix)
        print(name, os.path.getsize(os.path.join(OUT, name)) // 1024, "KiB", "loss", float(fix["loss"]))


if __name__ == "__main__":
    main()
