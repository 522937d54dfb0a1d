"""Times lsthm_gemm3 (tcgen05 split-bf16) against torch's fp32 SGEMM on the path's time-parallel shapes."""
import os
import sys
from importlib import import_module

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lsthm_b200  # noqa: E402

lib = import_module(lsthm_b200.__name__ + "._lib")
TN = 110 * 1024
SHAPES = [("NT gate_in v   x[TN,512] W[256,512]", lib.GEMM_NT, TN, 256, 512),
          ("NT gate_in l   x[TN,100] W[512,100]", lib.GEMM_NT, TN, 512, 100),
          ("NT enc qkv     x[TN,512] W[320,512]", lib.GEMM_NT, TN, 320, 512),
          ("NN dx          dy[TN,832] W[832,416]", lib.GEMM_NN, TN, 416, 832),
          ("TN dUV         dgx[TN,832] hz[TN,416]", lib.GEMM_TN, 832, 416, TN),
          ("TN dWatt       de[TN,832] c[TN,208]", lib.GEMM_TN, 832, 208, TN),
          ("TN dW enc      dy[TN,320] x[TN,512]", lib.GEMM_TN, 320, 512, TN),
          ("TN dWf2        dz[TN,208] u[TN,64]", lib.GEMM_TN, 208, 64, TN)]


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for name, mode, M, N, K in SHAPES:
    if mode == lib.GEMM_NT:
        a, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
        ref = lambda: a @ b.t()
    elif mode == lib.GEMM_NN:
        a, b = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
        ref = lambda: a @ b
    else:
        a, b = torch.randn(K, M, device="cuda"), torch.randn(K, N, device="cuda")
        ref = lambda: a.t() @ b
    t_ours, t_ref = timeit(lambda: lib.gemm3(mode, a, b)), timeit(ref)
    fl = 2.0 * M * N * K
    err = ((lib.gemm3(mode, a, b).double() - ref().double()).abs().max() / ref().abs().max()).item()
    print(f"{name:42s} ours {t_ours * 1e3:8.1f} us ({fl / t_ours / 1e9:6.1f} TF fp32-equiv)   torch fp32 {t_ref * 1e3:8.1f} us "
          f"({fl / t_ref / 1e9:5.1f} TF)   x{t_ref / t_ours:4.2f}   diff {err:.1e}")
