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


def test_every_kernel_binding_refuses_host_tensors():
    """No CPU fallback anywhere on the product path: each binding raises on host tensors before touching the library."""
    from importlib import import_module

    import lsthm_b200
    lib = import_module(lsthm_b200.__name__ + "._lib")
    fd = import_module(lsthm_b200.__name__ + ".fused_dln")
    pl = import_module(lsthm_b200.__name__ + ".pipeline")
    fa = import_module(lsthm_b200.__name__ + ".fused_attention")
    a, b = torch.randn(8, 8), torch.randn(8, 8)
    with pytest.raises(RuntimeError):
        lib.gemm3(lib.GEMM_NT, a, b)
    with pytest.raises(RuntimeError):
        lib.colsum(a)
    with pytest.raises(RuntimeError):
        fd.drop_res_layer_norm(a, None, b, torch.ones(8), torch.zeros(8), 1e-6)
    with pytest.raises(RuntimeError):
        pl.assemble_input(*(torch.zeros(2, 2, 8) for _ in range(5)))
    with pytest.raises(RuntimeError):
        fa.fused_self_attention(torch.randn(1, 4, 960), 8, 40 ** -0.5)
    with pytest.raises(RuntimeError):
        lib.adam_step(a.view(-1), b.view(-1), a.view(-1).clone(), a.view(-1).clone(), 1e-3, 0.9, 0.999, 1e-8, 0.0, 1)


def test_precision_switch_validates_and_sets_the_descriptor():
    from importlib import import_module

    import lsthm_b200
    lib = import_module(lsthm_b200.__name__ + "._lib")
    with pytest.raises(ValueError):
        lib.set_precision("fp16")
    try:
        lib.set_precision("bf16")
        assert lib.make_attn_desc(2, 5, 8, 960, 960, 960, 320, 0.1).precision == 1
        assert lib.make_attn_desc(2, 5, 8, 960, 960, 960, 320, 0.1, time_major=True).row_stride_i == 2
    finally:
        lib.set_precision("fp32")
    assert lib.make_attn_desc(2, 5, 8, 960, 960, 960, 320, 0.1).precision == 0


def test_rows_view_finds_copy_free_row_matrices():
    from importlib import import_module

    import lsthm_b200
    mm3 = import_module(lsthm_b200.__name__ + ".mm3")
    x = torch.randn(7, 3, 24)                                  # time-major storage [L, B, K]
    rows, swapped = mm3.rows_view(x)
    assert not swapped and rows.data_ptr() == x.data_ptr() and rows.shape == (21, 24)
    sl = x[:, :, 4:12].permute(1, 0, 2)                        # the view the reference hands to its encoders
    rows, swapped = mm3.rows_view(sl)
    assert swapped and rows.shape == (21, 8) and rows.stride() == (24, 1)
    assert torch.equal(rows.view(7, 3, 8), x[:, :, 4:12])
    assert mm3.rows_view(x.permute(2, 0, 1)) is None           # no unit inner stride -> caller copies
