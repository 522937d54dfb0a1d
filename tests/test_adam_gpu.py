"""GPU suite for the fused Adam step (lsthm_adam_step) and the flat-bucket optimizer: same trajectory as
torch.optim.Adam(lr=1e-3, weight_decay=2e-5) — the optimizer of the reference's trainer (model_trainer.py:82) —
including leaving never-used parameters untouched."""
from importlib import import_module

import pytest
import torch

import lsthm_b200
from helpers import masked_ce, seeded_model

pytestmark = pytest.mark.gpu
lib = import_module(lsthm_b200.__name__ + "._lib")
ddp = import_module(lsthm_b200.__name__ + ".ddp")


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 70001])
def test_flat_step_matches_torch_adam(n):
    g = torch.Generator(device="cuda").manual_seed(n)
    p = torch.randn(n, device="cuda", generator=g)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=2e-5)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        grad = torch.randn(n, device="cuda", generator=g)
        ref.grad = grad.clone()
        opt.step()
        lib.adam_step(p, grad, m, v, 1e-3, 0.9, 0.999, 1e-8, 2e-5, step)
        assert (p - ref.detach()).abs().max().item() <= 2e-7 * max(1.0, ref.abs().max().item())


def test_fused_adam_on_model_matches_torch():
    T, N = 6, 5
    a = seeded_model("AT", 41, "cuda").eval()
    b = seeded_model("AT", 41, "cuda").eval()
    red = ddp.GradAllReducer(a, 1, bucket_bytes=512 << 10, flatten_params=True)
    fused = ddp.FusedAdam(red, lr=1e-3, weight_decay=2e-5)
    opt = torch.optim.Adam(b.parameters(), lr=1e-3, weight_decay=2e-5)
    g = torch.Generator(device="cuda").manual_seed(2)
    for _ in range(3):
        x = torch.randn(T, N, 200, device="cuda", generator=g)
        lab = torch.randint(0, 7, (T * N,), device="cuda", generator=g)
        fused.zero_grad()
        masked_ce(a(x), lab, T, N).backward()
        red.finish()
        fused.step()
        opt.zero_grad(set_to_none=True)
        masked_ce(b(x), lab, T, N).backward()
        opt.step()
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert (pa - pb).abs().max().item() <= 1e-5 * max(1.0, pb.abs().max().item()), n
    assert a.encoder_l.pos_ffn.fc.weight.grad is None          # never used: stays out of the buckets
