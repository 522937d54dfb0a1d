"""GPU suite for the fused LayerNorm(dropout(y + bias) + residual) kernels (lsthm_dln_fwd/bwd) against an fp64
evaluation of model/encoder.py:54-58 / :106-112, row-strided operands, and forward/backward dropout consistency."""
from importlib import import_module

import pytest
import torch
import torch.nn.functional as F

import lsthm_b200

pytestmark = pytest.mark.gpu
fd = import_module(lsthm_b200.__name__ + ".fused_dln")


def _rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("R,d,with_bias,strided", [(1000, 100, True, False), (777, 512, False, True), (5, 200, True, True),
                                                   (1, 4, False, False), (4099, 100, False, True)])
def test_eval_forward_backward_vs_fp64(R, d, with_bias, strided):
    g = torch.Generator(device="cuda").manual_seed(R + d)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    y = rnd(R, d).requires_grad_(True)
    big = rnd(R, d + 12) if strided else None
    if strided:
        leaf = big.clone().requires_grad_(True)
        res_in = leaf[:, 4:4 + d]
    else:
        res_in = rnd(R, d).requires_grad_(True)
    gamma, beta = (1 + 0.1 * rnd(d)).requires_grad_(True), (0.1 * rnd(d)).requires_grad_(True)
    bias = (0.1 * rnd(d)).requires_grad_(True) if with_bias else None
    w = rnd(R, d)
    out = fd.drop_res_layer_norm(y, bias, res_in, gamma, beta, 1e-6)
    (out * w).sum().backward()
    y64, g64, b64 = (t.detach().double().requires_grad_(True) for t in (y, gamma, beta))
    r64 = res_in.detach().double().requires_grad_(True)
    bi64 = bias.detach().double().requires_grad_(True) if with_bias else None
    ref = F.layer_norm((y64 + bi64 if with_bias else y64) + r64, (d,), g64, b64, 1e-6)
    (ref * w.double()).sum().backward()
    assert _rel(out, ref.detach()) < 1e-5
    assert _rel(y.grad, y64.grad) < 2e-5
    rg = leaf.grad[:, 4:4 + d] if strided else res_in.grad
    assert _rel(rg, r64.grad) < 2e-5
    assert _rel(gamma.grad, g64.grad) < 2e-5 and _rel(beta.grad, b64.grad) < 2e-5
    if with_bias:
        assert _rel(bias.grad, bi64.grad) < 2e-5


def test_dropout_rate_seed_and_backward_mask():
    R, d, p = 4096, 100, 0.1
    g = torch.Generator(device="cuda").manual_seed(3)
    y = torch.randn(R, d, device="cuda", generator=g)
    res = torch.zeros(R, d, device="cuda")
    gamma, beta = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
    f = lambda yy, seed=9: fd.drop_res_layer_norm(yy, None, res, gamma, beta, 1e-6, p, seed)
    a, b, c = f(y), f(y), f(y, 10)
    assert torch.equal(a, b) and not torch.equal(a, c)
    # the mask itself: backward of sum(out * w) w.r.t. y is zero exactly on the dropped elements
    yy = y.clone().requires_grad_(True)
    w = torch.randn(R, d, device="cuda", generator=g)
    (f(yy) * w).sum().backward()
    dropped = (yy.grad == 0).float().mean().item()
    assert abs(dropped - p) < 0.005, dropped
    # directional derivative with the same seed
    dirn = torch.randn(R, d, device="cuda", generator=g)
    eps = 1e-3
    num = ((f(y + eps * dirn).double() * w).sum() - (f(y - eps * dirn).double() * w).sum()) / (2 * eps)
    ana = (yy.grad.double() * dirn).sum()
    assert abs(num - ana) / abs(ana) < 2e-2, (num.item(), ana.item())


def test_encoder_layer_fast_path_matches_module_path():
    """EncoderLayer on our kernels (2-D rows, time-major storage read in place) == the same module op by op."""
    enc_mod = import_module(lsthm_b200.__name__ + ".encoder")
    torch.manual_seed(0)
    L, B = 23, 5
    for d, d_inner in ((100, 50), (512, 50)):
        layer = enc_mod.EncoderLayer(d, d_inner, 8, 40, 40).cuda().eval()
        x_tm = torch.randn(L, B, d + 8, device="cuda")
        xa = x_tm.clone().requires_grad_(True)
        ya, _ = layer(xa[:, :, 4:4 + d].permute(1, 0, 2))
        w = torch.randn_like(ya)
        (ya * w).sum().backward()
        ga = {k: v.grad.clone() for k, v in layer.named_parameters() if v.grad is not None}
        layer.zero_grad()
        # fp64 truth: the module-by-module torch expressions (no kernel computes in fp64)
        xb = x_tm.clone().double().requires_grad_(True)
        layer64 = enc_mod.EncoderLayer(d, d_inner, 8, 40, 40).cuda().double().eval()
        layer64.load_state_dict({k: v.double() for k, v in layer.state_dict().items()})
        yb, _ = layer64(xb[:, :, 4:4 + d].permute(1, 0, 2))
        (yb * w.double()).sum().backward()
        assert _rel(ya, yb.detach()) < 2e-5
        assert _rel(xa.grad, xb.grad) < 1e-4
        for k, v in layer64.named_parameters():
            if v.grad is not None:
                assert _rel(ga[k], v.grad) < 1e-4, k


def test_encoder_layer_fast_path_train_mode_is_seeded_and_differentiates_its_own_mask():
    """Train mode of the fused layer (in-kernel dropout in the attention and in both LayerNorm tails, no mask tensors):
    the torch RNG seeds it, and the backward regenerates exactly the forward's masks (directional derivative)."""
    enc_mod = import_module(lsthm_b200.__name__ + ".encoder")
    torch.manual_seed(1)
    L, B, d = 31, 6, 100
    layer = enc_mod.EncoderLayer(d, 50, 8, 40, 40).cuda().train()
    x = torch.randn(L, B, d, device="cuda")
    w = torch.randn(B, L, d, device="cuda")

    def f(inp, seed):
        torch.manual_seed(seed)
        y, _ = layer(inp.permute(1, 0, 2))
        return (y * w).sum()

    a, b, c = f(x, 5), f(x, 5), f(x, 6)
    assert a.item() == b.item() and a.item() != c.item()
    layer.eval()
    assert f(x, 5).item() == f(x, 6).item() and f(x, 5).item() != a.item()      # eval: no dropout anywhere
    layer.train()
    xg = x.clone().requires_grad_(True)
    f(xg, 5).backward()
    dirn = torch.randn_like(x)
    eps = 1e-3
    num = (f(x + eps * dirn, 5).double() - f(x - eps * dirn, 5).double()) / (2 * eps)
    ana = (xg.grad.double() * dirn).sum()
    assert abs(num - ana) / abs(ana) < 2e-2, (num.item(), ana.item())
