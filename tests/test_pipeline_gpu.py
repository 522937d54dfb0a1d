"""GPU suite for the input side: fused batch assembly == the reference's ops bit for bit, and the
double-buffered feeder delivers the reference trainer's tensors (model_trainer.py:99-105) in order."""
from importlib import import_module

import pytest
import torch

import lsthm_b200

pytestmark = pytest.mark.gpu
pl = import_module(lsthm_b200.__name__ + ".pipeline")


@pytest.mark.parametrize("L,B", [(110, 32), (7, 3), (1, 1)])
def test_assemble_input_is_bit_identical(L, B):
    g = torch.Generator(device="cuda").manual_seed(L * B)
    r = [torch.randn(L, B, 1024, device="cuda", generator=g) for _ in range(4)]
    ac = torch.randn(L, B, 100, device="cuda", generator=g)
    ref = torch.cat(((r[0] + r[1] + r[2] + r[3]) / 4, ac), dim=-1)          # model_trainer.py:104-105
    assert torch.equal(pl.assemble_input(*r, ac), ref)


def test_feeder_yields_reference_batches_in_order():
    g = torch.Generator().manual_seed(0)
    batches = []
    for L, B in ((9, 4), (5, 2), (12, 3)):
        r = [torch.randn(L, B, 1024, generator=g) for _ in range(4)]
        vis, ac = torch.randn(L, B, 512, generator=g), torch.randn(L, B, 100, generator=g)
        qmask, umask = torch.zeros(L, B, 2), torch.ones(B, L)
        qmask[..., 0] = 1
        label = torch.randint(0, 6, (B, L), generator=g)
        batches.append((*r, vis, ac, qmask, umask, label, ["vid"] * B))
    got = list(pl.DeviceFeeder(batches, "cuda"))
    assert len(got) == 3
    for (x, qmask, umask, label), b in zip(got, batches):
        ref = torch.cat(((b[0] + b[1] + b[2] + b[3]) / 4, b[5]), dim=-1)
        assert x.is_cuda and torch.equal(x.cpu(), ref)
        assert torch.equal(qmask.cpu(), b[6]) and torch.equal(umask.cpu(), b[7]) and torch.equal(label.cpu(), b[8])


def test_host_tensors_are_rejected():
    t = torch.zeros(2, 2, 8)
    with pytest.raises(RuntimeError):
        pl.assemble_input(t, t, t, t, t)
