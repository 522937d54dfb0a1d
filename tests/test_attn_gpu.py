"""GPU suite for the fused encoder self-attention (tcgen05): forward and backward against an fp64 evaluation of
model/encoder.py:71-86 (softmax(q k^T / sqrt(d_k)) v, unmasked), ragged L <= 128, and consistency of the
in-kernel dropout between forward and backward."""
from importlib import import_module

import pytest
import torch

import lsthm_b200

pytestmark = pytest.mark.gpu
fa = import_module(lsthm_b200.__name__ + ".fused_attention")
H, D = 8, 40


def ref_attention(qkv, scale):
    B, L, W = qkv.shape
    q, k, v = (t.view(B, L, H, D).transpose(1, 2) for t in qkv.double().split(W // 3, dim=-1))
    p = torch.softmax((q * scale) @ k.transpose(-2, -1), dim=-1)
    return (p @ v).transpose(1, 2).reshape(B, L, H * D)


@pytest.mark.parametrize("B,L", [(2, 110), (3, 128), (5, 17), (1, 1), (4, 64)])
def test_forward_backward_vs_fp64(B, L):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + L)
    qkv = torch.randn(B, L, 3 * H * D, device="cuda", generator=g).requires_grad_(True)
    w = torch.randn(B, L, H * D, device="cuda", generator=g)
    scale = D ** -0.5
    out = fa.fused_self_attention(qkv, H, scale)
    (out * w).sum().backward()
    q64 = qkv.detach().double().requires_grad_(True)
    ref = ref_attention(q64, scale)
    (ref * w.double()).sum().backward()
    eo = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    eg = ((qkv.grad.double() - q64.grad).abs().max() / q64.grad.abs().max()).item()
    assert eo < 2e-5 and eg < 5e-5, (eo, eg)


def test_time_major_layout_matches_batch_major():
    """[L, B, .] activations (the reference's own layout) are read in place through the row strides."""
    B, L = 3, 37
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.randn(B, L, 3 * H * D, device="cuda", generator=g)
    w = torch.randn(B, L, H * D, device="cuda", generator=g)
    scale = D ** -0.5
    a = qkv.clone().requires_grad_(True)
    oa = fa.fused_self_attention(a, H, scale, 0.1, 3)
    (oa * w).sum().backward()
    bt = qkv.transpose(0, 1).contiguous().requires_grad_(True)
    ob = fa.fused_self_attention(bt, H, scale, 0.1, 3, time_major=True)
    (ob * w.transpose(0, 1)).sum().backward()
    assert torch.equal(oa, ob.transpose(0, 1)) and torch.equal(a.grad, bt.grad.transpose(0, 1))


def test_dropout_is_seeded_and_consistent_between_fwd_and_bwd():
    B, L, p = 2, 50, 0.1
    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = torch.randn(B, L, 3 * H * D, device="cuda", generator=g)
    w = torch.randn(B, L, H * D, device="cuda", generator=g)
    scale = D ** -0.5
    f = lambda x, seed=5: (fa.fused_self_attention(x, H, scale, p, seed) * w).sum()
    a, b, c = f(qkv), f(qkv), f(qkv, 6)
    assert a.item() == b.item() and a.item() != c.item()
    # keep rate ~ 1 - p: with v = 1 the output is sum_j mask_ij p_ij / (1-p), whose mean over rows is ~1
    ones = qkv.clone(); ones[:, :, 2 * H * D:] = 1.0
    m = fa.fused_self_attention(ones, H, scale, p, 11).mean().item()
    assert abs(m - 1.0) < 0.02, m
    # directional derivative: the backward must use the same mask as the forward
    x = qkv.clone().requires_grad_(True)
    f(x).backward()
    d = torch.randn(qkv.shape, device="cuda", generator=g)
    eps = 1e-2
    num = (f(qkv + eps * d).double() - f(qkv - eps * d).double()) / (2 * eps)
    ana = (x.grad.double() * d.double()).sum()
    assert abs(num - ana) / abs(ana) < 2e-2, (num.item(), ana.item())
