"""Kernel-only timing of the speaker-state cells at the benchmark shape (GPU box): python profiles/dev/cell_time.py"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
from importlib import import_module
import torch
import lsthm_b200
from helpers import sps_seeded_model
rec = import_module(lsthm_b200.__name__ + ".recurrence")
T, N = 110, 1024
g = torch.Generator().manual_seed(0)
q = torch.zeros(T, N, 2); s = torch.randint(0, 2, (N,), generator=g)
for t in range(T):
    s = torch.where(torch.rand(N, generator=g) < 0.6, 1 - s, s); q[t, torch.arange(N), s] = 1
for kind, rows in (("onlysp", 0), ("onlysp", 7), ("sps", 0), ("sps", 7)):
    model = sps_seeded_model(1, True, "cuda", kind=kind).train()
    cell = model.marn_cell_f
    cell.rows_per_cta = rows          # 0: planned tile (256 threads, <= 4 dialogues, two CTAs per SM); 7: 512-thread variant
    x_l, x_a, u = (torch.randn(T, N, d, device="cuda", requires_grad=True) for d in (100, 100, 200))
    qm = q.cuda()
    def step():
        out = cell(None if kind == "sps" else u, x_l, x_a, qm)
        out.sum().backward()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    rec.kernel_events = {"fwd": [], "bwd": []}
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    print(kind, "rows_per_cta", rows, {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v), 3) for k, v in rec.kernel_events.items()}, "ms per launch")
    rec.kernel_events = None
