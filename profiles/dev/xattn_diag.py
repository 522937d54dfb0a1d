"""Diagnostic (GPU box): which part of the sequence-level attention path carries the error of the perturbed-weights module
cases — projections (six-term GEMM) or core (lsthm_xattn) — by swapping each for its fp32 torch expression."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from importlib import import_module
import numpy as np
import torch
import lsthm_b200
from helpers import e_inf, sps_port_run, sps_port_run64, sps_run_module, load_golden, GOLDEN
sa = import_module(lsthm_b200.__name__ + ".seq_attention")

def torch_core(q, kv, B, L, scale, p, seed):
    D = q.shape[1]
    q3, k3, v3 = (t.view(L, B, -1).permute(1, 0, 2) for t in (q, kv[:, :D], kv[:, D:]))
    w = torch.softmax((q3 * scale) @ k3.transpose(1, 2), dim=-1)
    return (w @ v3).permute(1, 0, 2).reshape(L * B, D)

def dialogues(T, N, g):
    q = torch.zeros(T, N, 2); s = torch.randint(0, 2, (N,), generator=g)
    for t in range(T):
        flip = torch.rand(N, generator=g) < 0.6
        s = torch.where(flip, 1 - s, s); q[t, torch.arange(N), s] = 1
    q[3] = torch.tensor([1.0, 0.0]); q[4] = torch.tensor([0.0, 1.0])
    return q

T, N, seed = 10, 64, 123
g = torch.Generator().manual_seed(seed)
fix = {"seed": np.array(seed), "perturb": np.array(1), "train": np.array(0), "sample_stride": np.array(97),
       "x": torch.randn(T, N, 1124, generator=g).numpy(), "qmask": dialogues(T, N, g).numpy(),
       "umask": np.ones((N, T), np.float32), "labels": torch.randint(0, 6, (N, T), generator=g).numpy()}
cases = {"shard64": fix, "onlysp_s122": load_golden(os.path.join(GOLDEN, "onlysp_s122_T8_N7_eval_pert.npz")),
         "sps_s116": load_golden(os.path.join(GOLDEN, "sps_s116_T8_N7_eval_pert.npz"))}
orig_proj, orig_core = sa._ProjX6.apply, sa._SeqAttnCore.apply
for name, fx in cases.items():
    lp32, _, dx32, _ = sps_port_run(fx)
    truth, _ = sps_port_run64(fx)
    print(name, "ref32 vs 64: probs %.2e dx %.2e" % (e_inf(lp32, truth["probs64"]), e_inf(dx32, truth["dx64"])))
    for label, proj, core in (("x6+xattn", orig_proj, orig_core), ("torch+xattn", lambda x, w: x @ w, orig_core),
                              ("x6+torch", orig_proj, torch_core), ("torch+torch", lambda x, w: x @ w, torch_core)):
        sa._ProjX6.apply, sa._SeqAttnCore.apply = staticmethod(proj), staticmethod(core)
        logp, loss, dx, grads = sps_run_module(fx, rows_per_cta=7 if name == "shard64" else 0)
        worst = ("", 0.0, 0.0)
        if any(k.startswith("gsamp64/") for k in fx):
            stride = int(fx["sample_stride"])
            for k in fx:
                if not k.startswith("gsamp64/") or grads.get(k[8:]) is None:
                    continue
                flat = np.asarray(grads[k[8:]]).reshape(-1)
                samp = flat if flat.size <= 4096 else flat[::stride]
                bar = max(1e-3, 3.0 * e_inf(fx["gsamp/" + k[8:]], fx[k]))
                err = e_inf(samp, fx[k])
                if err / bar > worst[1]:
                    worst = (k[8:], err / bar, err)
        print("   %-12s probs %.2e dx %.2e   worst grad/bar %.2f (%s, err %.2e)" % (label, e_inf(logp, truth["probs64"]), e_inf(dx, truth["dx64"]), worst[1], worst[0], worst[2]))
sa._ProjX6.apply, sa._SeqAttnCore.apply = orig_proj, orig_core
