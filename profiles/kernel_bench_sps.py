"""Kernel-only timing of the speaker-state cell (one direction) through the C ABI.
Usage: python profiles/kernel_bench_sps.py [--N 1024] [--T 110] [--rows 0] [--reps 5] [--train]"""
import argparse
import os
import sys
from importlib import import_module

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lsthm_b200  # noqa: E402

lib = import_module(lsthm_b200.__name__ + "._lib")
sr = import_module(lsthm_b200.__name__ + ".sps_recurrence")

FLOP_FWD = 524_288 + 786_432 + 82_432          # speaker LSTMs + LSTHM1 (U,V,S) + collapsed attention, per utterance
FLOP_BWD = 524_288 + 786_432 + 3 * 82_432


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--T", type=int, default=110)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--train", action="store_true")
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    T, N = a.T, a.N
    cell = lsthm_b200.lsthm_sps.MARN1_sps(6).to(dev).marn_cell_f
    w = [t.detach().contiguous() for t in cell.cell_weights()]
    g = torch.Generator().manual_seed(1)
    s = torch.randint(0, 2, (N,), generator=g)
    q = torch.zeros(T, N, 2)
    for t in range(T):
        s = torch.where(torch.rand(N, generator=g) < 0.6, 1 - s, s)
        q[t, torch.arange(N), s] = 1
    qmask = q.to(dev)
    pi, pr, n0 = sr.party_plan(qmask)
    new = lambda *sh: torch.empty(*sh, device=dev)
    gx, dout = torch.randn(T, N, 2, 512, device=dev), torch.randn(T, N, 512, device=dev)
    masks = [None] * 5
    att_p = 0.0
    if a.train:
        masks = [(torch.rand(T, N, 128, device=dev) < 0.5).float() * 2 for _ in range(4)] + [None]
        att_p = 0.2
    desc = lib.make_sps_desc(T, N, a.rows, att_p, 1234)
    ws_ = lib.make_sps_weights(w[0:2], w[2:4], w[4:6], w[6:8], w[8:10], w[10:12], w[12], w[13])
    mk = lib.make_sps_masks(*masks)
    packed = new(lib.sps_packed_floats())
    lib.sps_pack(ws_, packed)
    work = new(lib.sps_workspace_floats(desc))
    out = new(T, N, 512)
    sGQ, sGL = new(T, N, 2, 512), new(T, N, 2, 512)
    sCQ, sHQ, sXQ, sCL, sHL = (new(T, N, 2, 128) for _ in range(5))
    dGL, dGQ = new(T, N, 2, 512), new(T, N, 2, 512)
    info = lib.sps_launch_info(desc)
    dWqk = new(info["grid"], 2, 128)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tf, tb = [], []
    for i in range(a.reps + 2):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        lib.sps_fwd(desc, ws_, packed, gx, qmask, pi, n0, mk, work, out, sGQ, sCQ, sHQ, sXQ, sGL, sCL, sHL)
        e1.record()
        lib.sps_bwd(desc, ws_, qmask, pi, pr, n0, mk, dout, sGQ, sCQ, sGL, sCL, work, dGL, dGQ, dWqk)
        e2.record()
        torch.cuda.synchronize()
        if i >= 2:
            tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
    f, b = min(tf), min(tb)
    print(f"sps cell (one direction) N={N} T={T} rows={info['rows']} grid={info['grid']} train={a.train}: "
          f"fwd {f:.3f} ms ({FLOP_FWD * T * N / f / 1e9:.1f} TFLOP/s)  bwd {b:.3f} ms ({FLOP_BWD * T * N / b / 1e9:.1f} TFLOP/s)")


if __name__ == "__main__":
    main()
