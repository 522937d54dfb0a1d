"""GPU suite for the sequence-level cross-modal attention kernels (model/lsthm_sps.py:88-129) through the
C ABI: ``lsthm_xattn_fwd/bwd`` (tcgen05 attention core, width 128 / 100) and the six-term GEMM mode ``LSTHM_GEMM_X6``.

  1. core vs an fp64 torch restatement: out, dq, dk, dv for several (B, L, D), ragged L, the benchmark shape;
  2. in-kernel dropout: seeded, forward and backward use the same mask (directional derivative), rate and scaling;
  3. six-term projection of a zero-sum (LayerNorm-output) operand by all-ones weights — the case the three-term split cannot
     do — lands at fp32 rounding level;
  4. the module-level function (projections + core) vs fp64, ones-initialised AND perturbed weights.
"""
from importlib import import_module

import pytest
import torch

import lsthm_b200
from helpers import e_inf

pytestmark = pytest.mark.gpu
_lib = import_module(lsthm_b200.__name__ + "._lib")
sa = import_module(lsthm_b200.__name__ + ".seq_attention")


def _ref(q, k, v, B, L, scale):
    """fp64 restatement on time-major rows (row = i*B + b)."""
    D = q.shape[1]
    q3, k3, v3 = (t.double().view(L, B, -1).permute(1, 0, 2) for t in (q, k, v))
    w = torch.softmax((q3 * scale) @ k3.transpose(1, 2), dim=-1)
    return (w @ v3).permute(1, 0, 2).reshape(L * B, D)


@pytest.mark.parametrize("B,L,D", [(3, 9, 128), (5, 110, 128), (2, 128, 128), (4, 37, 100), (7, 110, 100), (1, 2, 8), (64, 110, 128)])
def test_core_vs_fp64(B, L, D):
    g = torch.Generator().manual_seed(B * 1000 + L)
    q = torch.randn(L * B, D, generator=g) * 0.7
    kv = torch.randn(L * B, 2 * D, generator=g)
    dout = torch.randn(L * B, D, generator=g)
    scale = 1.0 / D ** 0.5
    q64, kv64 = q.double().requires_grad_(True), kv.double().requires_grad_(True)
    ref = _ref(q64, kv64[:, :D], kv64[:, D:], B, L, scale)
    (ref * dout.double()).sum().backward()
    qc, kvc = q.cuda().requires_grad_(True), kv.cuda().requires_grad_(True)
    out = sa._SeqAttnCore.apply(qc, kvc, B, L, scale, 0.0, 0)
    (out * dout.cuda()).sum().backward()
    errs = {"out": e_inf(out.detach().cpu(), ref.detach()), "dq": e_inf(qc.grad.cpu(), q64.grad), "dkv": e_inf(kvc.grad.cpu(), kv64.grad)}
    assert errs["out"] <= 2e-5 and errs["dq"] <= 1e-4 and errs["dkv"] <= 1e-4, errs


def test_core_dropout_is_seeded_and_consistent():
    B, L, D, p = 6, 110, 128, 0.2
    g = torch.Generator().manual_seed(5)
    q, kv = (torch.randn(L * B, D, generator=g) * 0.5).cuda(), torch.randn(L * B, 2 * D, generator=g).cuda()
    scale = 1.0 / D ** 0.5
    f = lambda q_, kv_, seed: sa._SeqAttnCore.apply(q_, kv_, B, L, scale, p, seed)
    a, b, c = f(q, kv, 11), f(q, kv, 11), f(q, kv, 12)
    assert torch.equal(a, b) and not torch.equal(a, c)
    # with V = 1 the output is the kept probability mass / (1-p): mean 1, and a fraction ~p of the L entries per row is dropped
    ones = kv.clone(); ones[:, D:] = 1.0
    mass = f(q, ones, 3)[:, 0]
    assert abs(float(mass.mean()) - 1.0) < 0.02 and float(mass.std()) > 0.01
    # forward and backward evaluate the same mask: directional derivative of sum(out * w) along (dq, dkv)
    qg, kvg = q.clone().requires_grad_(True), kv.clone().requires_grad_(True)
    w = torch.randn(L * B, D, generator=g).cuda()
    (f(qg, kvg, 7) * w).sum().backward()
    dq, dkv = torch.randn(q.shape, generator=g).cuda(), torch.randn(kv.shape, generator=g).cuda()
    eps = 1e-2
    num = ((f(q + eps * dq, kv + eps * dkv, 7) * w).double().sum() - (f(q - eps * dq, kv - eps * dkv, 7) * w).double().sum()) / (2 * eps)
    ana = (qg.grad * dq).double().sum() + (kvg.grad * dkv).double().sum()
    assert abs(float(num - ana)) <= 2e-3 * max(abs(float(ana)), 1.0), (float(num), float(ana))


def test_six_term_projection_of_cancelling_sums():
    """LayerNorm output (zero row sums up to rounding) times all-ones weights: the exact answer is rounding noise."""
    g = torch.Generator().manual_seed(2)
    M, K, N = 4096, 100, 128
    x = torch.nn.functional.layer_norm(torch.randn(M, K, generator=g), (K,)).cuda()
    w = torch.ones(K, N).cuda()
    truth = (x.double() @ w.double())
    y6 = _lib.gemm3(_lib.GEMM_NN, x, w, x6=True)
    y3 = _lib.gemm3(_lib.GEMM_NN, x, w)
    y32 = x @ w                                                   # library fp32 product, for scale only
    err = lambda y: float((y.double() - truth).abs().max())
    assert err(y6) <= 4 * max(err(y32), 1e-6), (err(y6), err(y32))
    assert err(y3) > 10 * err(y6), (err(y3), err(y6))            # documents why three terms are not enough here
    # and a generic product: six terms at fp32 level
    a, b = torch.randn(M, K, generator=g).cuda(), torch.randn(K, N, generator=g).cuda()
    t = a.double() @ b.double()
    assert e_inf(_lib.gemm3(_lib.GEMM_NN, a, b, x6=True).cpu(), t.cpu()) < 2e-6


@pytest.mark.parametrize("ones", [True, False])
@pytest.mark.parametrize("d2,dk", [(100, 128), (128, 128), (100, 100)])
def test_projections_plus_core_vs_fp64(ones, d2, dk):
    L, B, d1 = 23, 5, 100
    g = torch.Generator().manual_seed(9)
    ln = lambda t: torch.nn.functional.layer_norm(t, (t.shape[-1],))
    x1, x2 = ln(torch.randn(L, B, d1, generator=g)), ln(torch.randn(L, B, d2, generator=g))
    mk = (lambda *s: torch.ones(*s)) if ones else (lambda *s: 1.0 + 0.1 * torch.randn(*s, generator=g))
    Wq, Wk, Wv = mk(d1, dk), mk(d2, dk), mk(d2, dk)
    dout = torch.randn(L, B, dk, generator=g)

    def run(dt, dev):
        ts = [t.to(dev, dt).requires_grad_(True) for t in (x1, x2, Wq, Wk, Wv)]
        a, b = ts[0].permute(1, 0, 2), ts[1].permute(1, 0, 2)
        if dev == "cuda":
            out = sa.seq_cross_attention(ts[0], ts[1], ts[2], ts[3], ts[4])
        else:
            w = torch.softmax(((a @ ts[2]) / dk ** 0.5) @ (b @ ts[3]).transpose(1, 2), dim=-1)
            out = (w @ (b @ ts[4])).permute(1, 0, 2)
        (out * dout.to(dev, dt)).sum().backward()
        return [out.detach().cpu()] + [t.grad.cpu() for t in ts]
    ours, truth, ref32 = run(torch.float32, "cuda"), run(torch.float64, "cpu"), run(torch.float32, "cpu")
    for name, o, t, r in zip(("out", "dx1", "dx2", "dWq", "dWk", "dWv"), ours, truth, ref32):
        scale = float(t.abs().max())
        bar = max(1e-4 if name == "out" else 1e-3, 3.0 * e_inf(r, t)) if scale > 1e-9 else None
        if bar is None:
            continue
        # absolute floor: with ones weights Q, K, V are rounding noise (1e-6) and so are some gradients; compare at the scale of dout
        assert float((o.double() - t).abs().max()) <= bar * max(scale, 1e-4), (name, e_inf(o, t), bar, scale)
