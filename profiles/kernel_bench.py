"""Kernel-only timing of the recurrence forward/backward through the C ABI (CUDA events on the
launching stream, inputs resident, L2 flushed between launches by the 1.4 GB of stash traffic).
Usage: python profiles/kernel_bench.py [--N 1024] [--T 110] [--rows 0] [--reps 5] [--kind ATV]"""
import argparse
import os
import sys
from importlib import import_module

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lsthm_b200  # noqa: E402

lib = import_module(lsthm_b200.__name__ + "._lib")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--T", type=int, default=110)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--kind", default="ATV")
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    model = (lsthm_b200.HybridRNN_ATV.MARN() if a.kind == "ATV" else lsthm_b200.HybridRNN_AT.MARN()).to(dev)
    dh, rd, MH = model._dh, model._rd, 64
    D, R, M = sum(dh), sum(rd), len(dh)
    T, N = a.T, a.N
    w = [t.detach().contiguous() for t in model.recurrence_weights()]
    d = lib.make_desc(T, N, dh, rd, MH, 4, a.rows)
    ws = lib.make_weights(w[0:M], w[M:2 * M], w[2 * M], w[2 * M + 1], w[2 * M + 2:3 * M + 2], w[3 * M + 2:4 * M + 2],
                          *w[4 * M + 2:4 * M + 6])
    packed = torch.empty(lib.mab_packed_floats(d), device=dev)
    lib.mab_pack(d, ws, packed)
    new = lambda *s: torch.empty(*s, device=dev)
    gx, dhz = torch.randn(T, N, 4 * D, device=dev), torch.randn(T, N, 2 * D, device=dev)
    mask = (torch.rand(T, N, MH, device=dev) < 0.7).float() / 0.7
    hz, sC, sG, sA, sU = new(T, N, 2 * D), new(T, N, D), new(T, N, 4 * D), new(T, N, 4 * D), new(T, N, MH)
    dgx, de, dup, att, duz = new(T, N, 4 * D), new(T, N, 4 * D), new(T, N, MH), new(T, N, 4 * D), torch.randn(T, N, MH, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tf, tb = [], []
    for i in range(a.reps + 2):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        lib.mab_fwd(d, packed, gx, mask, hz, sU, sC, sG, sA)
        e1.record()
        lib.mab_bwd(d, ws, packed, dhz, duz, mask, sC, sG, sA, sU, dgx, de, dup, att)
        e2.record()
        torch.cuda.synchronize()
        if i >= 2:
            tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
    info = lib.mab_launch_info(d)
    flop = 999_936 * T * N if a.kind == "ATV" else 0
    f, b = min(tf), min(tb)
    print(f"{os.environ.get('LSTHM_B200_SO', 'default')}: kind={a.kind} N={N} T={T} rows={info['rows']} grid={info['grid']} "
          f"fwd {f:.3f} ms bwd {b:.3f} ms" + (f"  ({flop / f / 1e9:.1f} / {flop / b / 1e9:.1f} TFLOP/s)" if flop else ""))


if __name__ == "__main__":
    main()
