"""GPU suite for the tcgen05 split-bf16 GEMM (lsthm_gemm3) against an fp64 product: all three operand
layouts, ragged M/N/K (zero-filled tails), bias, split-K with the deterministic reduce."""
from importlib import import_module

import pytest
import torch

import lsthm_b200

pytestmark = pytest.mark.gpu
lib = import_module(lsthm_b200.__name__ + "._lib")


def _err(c, ref):
    return ((c.double() - ref).abs().max() / ref.abs().max()).item()


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 128, 64), (1000, 832, 100), (4100, 64, 512), (333, 208, 244),
                                   (128, 8, 16), (5000, 320, 100)])
def test_nt_linear_forward(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.1
    b = torch.randn(N, device="cuda", generator=g)
    ref = x.double() @ w.double().t() + b.double()
    assert _err(lib.gemm3(lib.GEMM_NT, x, w, b), ref) < 2e-5
    assert _err(lib.gemm3(lib.GEMM_NT, x, w), ref - b.double()) < 2e-5


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (3000, 100, 832), (777, 512, 256)])
def test_nn_input_gradient(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    dy = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(K, N, device="cuda", generator=g) * 0.1
    assert _err(lib.gemm3(lib.GEMM_NN, dy, w), dy.double() @ w.double()) < 2e-5


@pytest.mark.parametrize("M,N,K", [(128, 128, 4096), (832, 416, 20000), (512, 100, 9000), (64, 244, 33000), (16, 512, 5000)])
def test_tn_weight_gradient_split_k(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    dy = torch.randn(K, M, device="cuda", generator=g)
    x = torch.randn(K, N, device="cuda", generator=g)
    ref = dy.double().t() @ x.double()
    c0 = lib.gemm3(lib.GEMM_TN, dy, x)
    assert _err(c0, ref) < 2e-5
    assert torch.equal(c0, lib.gemm3(lib.GEMM_TN, dy, x))          # split-K reduce is order-fixed


def test_strided_views():
    """Operands that are column slices of wider tensors (how the recurrence adjoints are used)."""
    big = torch.randn(6000, 832, device="cuda")
    act = torch.randn(6000, 416, device="cuda")
    a, b = big[:, 512:576], act[:, 208:]
    assert _err(lib.gemm3(lib.GEMM_TN, a, b), a.double().t() @ b.double()) < 2e-5


@pytest.mark.parametrize("R,C,off,ld", [(112640, 832, 0, 832), (1000, 100, 4, 120), (7, 4, 0, 4), (513, 244, 16, 260), (65, 1024, 0, 1024)])
def test_colsum_matches_fp64_and_is_deterministic(R, C, off, ld):
    from importlib import import_module
    mm3 = import_module(lsthm_b200.__name__ + ".mm3")
    g = torch.Generator(device="cuda").manual_seed(R + C)
    big = torch.randn(R, ld, device="cuda", generator=g)
    a = big[:, off:off + C]
    s1, s2 = mm3.colsum(a), mm3.colsum(a)
    ref = a.double().sum(0)
    assert torch.equal(s1, s2)
    assert ((s1.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5


@pytest.mark.parametrize("M,N,K", [(4096, 960, 100), (3000, 100, 320), (2500, 52, 512), (2049, 244, 64), (9000, 512, 100),
                                   (2300, 8, 1024), (2100, 300, 36)])
def test_weight_stationary_path_nt_nn_relu(M, N, K):
    """Rows >= GEMM3W_MIN_ROWS route NT/NN through lsthm_gemm3w (pre-split weight images, 128 x 256 tiles): ragged
    N and K, N spanning several 256-column chunks, bias and the ReLU epilogue, strided activations."""
    assert M >= lib.GEMM3W_MIN_ROWS
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    xw = torch.randn(M, K + 8, device="cuda", generator=g)
    x = xw[:, 4:4 + K]
    w = torch.randn(N, K, device="cuda", generator=g) * 0.1
    b = torch.randn(N, device="cuda", generator=g)
    ref = x.double() @ w.double().t() + b.double()
    assert _err(lib.gemm3(lib.GEMM_NT, x, w, b), ref) < 2e-5
    assert _err(lib.gemm3(lib.GEMM_NT, x, w), ref - b.double()) < 2e-5
    assert _err(lib.gemm3(lib.GEMM_NT_RELU, x, w, b), ref.clamp_min(0)) < 2e-5
    dy = torch.randn(M, N, device="cuda", generator=g)
    assert _err(lib.gemm3(lib.GEMM_NN, dy, w), dy.double() @ w.double()) < 2e-5
