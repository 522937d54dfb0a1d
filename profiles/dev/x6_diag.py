"""Diagnostic (GPU box): absolute error of the six-term / three-term tensor-core GEMM and of the library fp32 SGEMM against fp64."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from importlib import import_module
import torch
import lsthm_b200
_lib = import_module(lsthm_b200.__name__ + "._lib")
g = torch.Generator().manual_seed(2)
M, K, N = 8192, 100, 128
ln = lambda t: torch.nn.functional.layer_norm(t, (t.shape[-1],))
cases = {"generic": (torch.randn(M, K, generator=g), torch.randn(K, N, generator=g)),
         "LN x ones": (ln(torch.randn(M, K, generator=g)), torch.ones(K, N)),
         "LN x (1+0.1n)": (ln(torch.randn(M, K, generator=g)), 1 + 0.1 * torch.randn(K, N, generator=g)),
         "LN x (1+0.1n), K=128": (ln(torch.randn(M, 128, generator=g)), 1 + 0.1 * torch.randn(128, N, generator=g))}
for name, (a, b) in cases.items():
    a, b = a.cuda(), b.cuda()
    t = a.double() @ b.double()
    err = lambda y: ((y.double() - t).abs().max().item(), (y.double() - t).pow(2).mean().sqrt().item(), (y.double() - t).mean().item())
    print(name, "|t|max %.3g" % t.abs().max().item())
    print("   x6    max %.2e rms %.2e mean %+.2e" % err(_lib.gemm3(_lib.GEMM_NN, a, b, x6=True)))
    print("   x3    max %.2e rms %.2e mean %+.2e" % err(_lib.gemm3(_lib.GEMM_NN, a, b)))
    print("   sgemm max %.2e rms %.2e mean %+.2e" % err(a @ b))
    print("   cpu   max %.2e rms %.2e mean %+.2e" % err((a.cpu() @ b.cpu()).cuda()))
