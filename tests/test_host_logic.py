"""CPU suite, part 2: the host side above the C ABI, with the CUDA library replaced by the plain-C
oracle (tests/oracle_backend.py).  Checks the drop-in modules end to end against the
reference-generated fixtures: module wiring, autograd.Function plumbing and the hoisted
weight-gradient products of recurrence.py.  Also: the product path refuses to run without CUDA."""
import pytest
import torch

import oracle_backend
from helpers import SPEC, check_against_golden, golden_files, load_golden, run_module, seeded_model


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_module_with_oracle_backend_matches_fixture(path, monkeypatch):
    oracle_backend.install(monkeypatch)
    fix = load_golden(path)
    probs, loss, dx, grads = run_module(fix)
    check_against_golden(fix, probs, loss, dx, grads, tol_out=5e-6, tol_grad=5e-5)


def test_product_path_refuses_cpu():
    model = seeded_model("AT", 1).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.randn(3, 2, SPEC["AT"]["din"]))


def test_unused_parameters_stay_gradless(monkeypatch):
    oracle_backend.install(monkeypatch)
    model = seeded_model("AT", 1).eval()
    model(torch.randn(3, 2, 200)).sum().backward()
    none = [n for n, p in model.named_parameters() if p.grad is None]
    assert sorted(none) == sorted(f"encoder_{m}.pos_ffn.fc.{w}" for m in "la" for w in ("weight", "bias"))
