"""Extra measurement (not part of bench.py's contract): the REFERENCE ALGORITHM as plain PyTorch eager ops on the same
B200 — the oracle's torch restatement of HybridRNN_ATV (oracle/torch_port.py, a per-step Python loop of library kernels,
exactly what the reference's own code does on a GPU; the reference tree itself cannot travel to the GPU box) — fwd+bwd
at the headline shape, next to our drop-in module.  Measurement infrastructure only: the product never imports oracle/.
    python profiles/eager_gpu_baseline.py [--batch 1024] [--seq 110] [--steps 3]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lsthm_b200  # noqa: E402
from oracle import torch_port as tp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--seq", type=int, default=110)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(111)
T, B = a.seq, a.batch
model = lsthm_b200.HybridRNN_ATV.MARN().to(dev).train()
params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
x = torch.randn(T, B, 712, device=dev)
labels = torch.randint(0, 6, (T * B,), device=dev)
umask = torch.ones(B, T, device=dev)


def eager_step():
    for p in params.values():
        p.grad = None
    probs = tp.mab_forward(params, x, "ATV", None)      # no dropout kernels at all (the tape lives on the host): favours the baseline
    tp.masked_loss(probs, labels, umask).backward()


def ours_step():
    model.zero_grad(set_to_none=True)
    tp.masked_loss(model(x), labels, umask).backward()


for name, fn in (("pytorch-eager restatement of the reference", eager_step), ("lsthm_b200 drop-in module", ours_step)):
    fn()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(a.steps):
        fn()
    torch.cuda.synchronize()
    ms = (time.time() - t0) / a.steps * 1e3
    print(f"{name}: {ms:.1f} ms per fwd+bwd step, {T * B / ms * 1e3:.0f} utterances/s  (ATV, {B} dialogues x {T}, fp32; ours in train mode, the eager baseline without dropout)")
