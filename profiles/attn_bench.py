"""Times lsthm_attn_fwd / lsthm_attn_bwd alone at the headline shape (B=1024 dialogues, L=110, 8 heads x 40) with CUDA
events on the launching stream, inputs alternating between two buffers larger than L2.
    python profiles/attn_bench.py [p_drop]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lsthm_b200  # noqa: E402
from importlib import import_module  # noqa: E402

lib = import_module(lsthm_b200.__name__ + "._lib")
B, L, H, p = 1024, 110, 8, float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
TM = (sys.argv[2] if len(sys.argv) > 2 else "tm") == "tm"      # time-major [L,B,.] rows (the layout the model uses) or batch-major
W, HD = 3 * H * 40, H * 40
S0, S1 = (L, B) if TM else (B, L)
bufs = [torch.randn(S0, S1, W, device="cuda") for _ in range(2)]
douts = [torch.randn(S0, S1, HD, device="cuda") for _ in range(2)]
out = torch.empty(S0, S1, HD, device="cuda")
lse = torch.empty(B * H, L, device="cuda")
dqkv = torch.empty(S0, S1, W, device="cuda")
d = lib.make_attn_desc(B, L, H, W, W, W, HD, 40 ** -0.5, p, 1234, time_major=TM)


def fwd(i):
    q = bufs[i & 1]
    lib.attn_fwd(d, q[:, :, :HD], q[:, :, HD:2 * HD], q[:, :, 2 * HD:], out, lse)


def bwd(i):
    q = bufs[i & 1]
    lib.attn_bwd(d, q[:, :, :HD], q[:, :, HD:2 * HD], q[:, :, 2 * HD:], out, lse, douts[i & 1], dqkv[:, :, :HD], dqkv[:, :, HD:2 * HD],
                 dqkv[:, :, 2 * HD:])


for name, fn, flops, nbytes in (("fwd", fwd, 4.0 * B * H * L * L * 40, 4 * (B * L * W + B * L * HD)),
                                ("bwd", bwd, 10.0 * B * H * L * L * 40, 4 * (2 * B * L * W + 2 * B * L * HD))):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"attn_{name}: {ms * 1e3:.1f} us/launch  {flops / ms / 1e9:.1f} TFLOP/s fp32-equivalent  {nbytes / ms / 1e6:.0f} GB/s algorithmic HBM (p_drop={p}, {'time' if TM else 'batch'}-major)")
