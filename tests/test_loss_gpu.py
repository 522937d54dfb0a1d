"""GPU test of the fused MaskedLoss kernels (loss.py:13-21 of the reference) through the C ABI: value and gradient against the
reference's own expression in fp64, cross entropy and NLL, ragged masks (padded rows contribute log C and no gradient), 6 / 7
classes, the benchmark size; bitwise run-to-run determinism; the weighted form stays on the reference's expression."""
import math

import pytest
import torch

import lsthm_b200

pytestmark = pytest.mark.gpu


def _reference(losser, pred, target, mask):
    flat = mask.reshape(-1, 1)
    return losser(reduction="sum")(pred * flat, target) / mask.sum()


@pytest.mark.parametrize("losser", [torch.nn.CrossEntropyLoss, torch.nn.NLLLoss])
@pytest.mark.parametrize("B,L,C", [(5, 9, 6), (3, 7, 7), (1024, 110, 6), (2, 1, 2)])
def test_masked_loss_matches_reference_expression(losser, B, L, C):
    g = torch.Generator().manual_seed(B * 10 + C)
    pred = torch.randn(B * L, C, generator=g)
    if losser is torch.nn.NLLLoss:
        pred = torch.log_softmax(pred, -1)
    target = torch.randint(0, C, (B * L,), generator=g)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    mask = (torch.arange(L)[None, :] < lens[:, None]).float()
    p64 = pred.double().requires_grad_(True)
    ref = _reference(losser, p64, target, mask.double())
    ref.backward()
    pc = pred.cuda().requires_grad_(True)
    crit = lsthm_b200.MaskedLoss(losser)
    out = crit(pc, target.cuda(), mask.cuda())
    (out * 1.7).backward()
    assert abs(float(out) - float(ref)) <= 2e-6 * max(1.0, abs(float(ref)))
    assert float((pc.grad.cpu().double() - 1.7 * p64.grad).abs().max()) <= 2e-6 * float(p64.grad.abs().max()) + 1e-12
    out2 = crit(pc.detach(), target.cuda(), mask.cuda())
    assert torch.equal(out.detach(), out2)                                   # deterministic
    if losser is torch.nn.CrossEntropyLoss and int((mask == 0).sum()) > 0:
        # a padded row is an all-zero logit row: log C each, no gradient
        rows = (mask.reshape(-1) == 0)
        assert float(pc.grad[rows.cuda()].abs().max()) == 0.0
        only_pad = _reference(losser, torch.zeros(int(rows.sum()), C), target[rows], torch.ones(int(rows.sum()))) 
        assert abs(float(only_pad) - math.log(C)) < 1e-6


def test_weighted_loss_keeps_the_reference_expression():
    w = torch.tensor([1.0, 2.0, 0.5, 1.0, 1.0, 3.0]).cuda()
    crit = lsthm_b200.MaskedLoss(torch.nn.CrossEntropyLoss, weight=w)
    g = torch.Generator().manual_seed(0)
    pred, target, mask = torch.randn(12, 6, generator=g).cuda(), torch.randint(0, 6, (12,), generator=g).cuda(), torch.ones(3, 4).cuda()
    flat = mask.reshape(-1, 1)
    want = torch.nn.CrossEntropyLoss(weight=w, reduction="sum")(pred * flat, target) / (w[target] * flat.squeeze(1)).sum()
    assert torch.allclose(crit(pred, target, mask), want)
