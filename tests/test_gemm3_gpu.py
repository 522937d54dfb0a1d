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
