"""Drop-in boundary under the reference's OWN training loop: the unmodified
``model_trainer.py`` (staged byte for byte in oracle/_ref) builds the model by name, owns Adam + StepLR, and runs
``train_network`` / ``eval_network``.  The only integration is that its ``from models.lsthm_<x> import MARN1_<x>`` lines
resolve to the mirror modules.  Checked against the same trainer driving the reference classes on the CPU:

  * ``eval_network``: identical predictions (accuracy, weighted F1) on a synthetic IEMOCAP-shaped loader;
  * ``train_network`` (dropout probabilities set to 0 on both sides so the two runs are comparable): the epoch loss over two
    optimizer steps, and the model's update vector after the two Adam steps (relative L2) — forward, backward and the
    optimizer through the reference's code path; parameters the reference never uses keep ``grad = None`` and are not decayed (F8).
"""
import os

import numpy as np
import pytest
import torch

import lsthm_b200
from oracle import ref_shim

pytestmark = pytest.mark.gpu


def _loader(seed, lens_per_batch):
    """Batches in the layout of IEMOCAPDataset.collate_fn (dataloader.py:29-47): r1..r4 [L,B,1024], visuf [L,B,512],
    acouf [L,B,100], qmask [L,B,2], umask [B,L], label [B,L], ids."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for lens in lens_per_batch:
        L, B = max(lens), len(lens)
        r = [torch.randn(L, B, 1024, generator=g) for _ in range(4)]
        visuf, acouf = torch.randn(L, B, 512, generator=g), torch.randn(L, B, 100, generator=g)
        qmask, umask, label = torch.zeros(L, B, 2), torch.zeros(B, L), torch.zeros(B, L, dtype=torch.long)
        for b, n in enumerate(lens):
            umask[b, :n] = 1
            label[b, :n] = torch.randint(0, 6, (n,), generator=g)
            s = int(torch.randint(0, 2, (1,), generator=g))
            for t in range(n):
                if t and torch.rand(1, generator=g).item() < 0.6:
                    s = 1 - s
                qmask[t, b, s] = 1
            for x in r + [visuf, acouf]:
                x[n:, b] = 0
        out.append(r + [visuf, acouf, qmask, umask, label, [f"d{b}" for b in range(B)]])
    return out


def _kink_free_perturbation(name, state, batches, thr=4e-6, tries=24):
    """Seed of ``perturb_ones`` for which no ReLU pre-activation of the model (fp64 oracle run on the test batches) lies within
    ``thr`` (relative to its layer's largest) of zero.  A unit that close to its kink is switched on or off by rounding
    noise — any two fp32 implementations may disagree about it, which moves every upstream gradient by the
    whole contribution of that unit (observed with seed 7 and MARN1_nsps: one unit of nn_out.0 at 2e-6, gradients 5-20 %
    apart while the loss agrees to the last digit).  The comparison below needs data that is decidable at parity resolution."""
    from helpers import relu_kinks
    from oracle import torch_port as tp
    fwd = {"MARN1_onlysp": tp.onlysp_forward, "MARN1_nsps": tp.nsps_forward, "MARN1_no_en": tp.no_en_forward,
           "MARN1_sps": tp.sps_forward}[name]
    for seed in range(7, 7 + tries):
        p = {k[len("model."):]: v.detach().clone() for k, v in state.items() if k.startswith("model.")}
        tp.perturb_ones(p, seed)
        p = {k: v.double() for k, v in p.items()}
        worst = 1.0
        for d in batches:
            x = torch.cat(((d[0] + d[1] + d[2] + d[3]) / 4, d[5]), -1).double()
            with relu_kinks() as rk:
                fwd(p, x, d[6].double(), d[7].double(), None)
            worst = min(worst, min(float(z.abs().min() / z.abs().max()) for z in rk.z))
        if worst >= thr:
            return seed
    pytest.skip("no kink-free perturbation seed found")


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference neither at /root/reference nor staged in oracle/_ref")
@pytest.mark.parametrize("name", ["MARN1_onlysp", "MARN1_nsps", "MARN1_no_en", "MARN1_sps"])
def test_reference_trainer_drives_the_dropin(name, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                                   # eval_network writes res.csv into the cwd (model_trainer.py:158)
    mod = {"MARN1_onlysp": "lsthm_onlysp", "MARN1_nsps": "lsthm_nsps", "MARN1_no_en": "lsthm_no_en", "MARN1_sps": "lsthm_sps"}[name]
    ref_mt = ref_shim.load_trainer(name="_mt_reference")
    our_mt = ref_shim.load_trainer({f"models.{mod}": getattr(lsthm_b200, mod)}, name="_mt_dropin")
    assert getattr(our_mt, name) is getattr(getattr(lsthm_b200, mod), name) and getattr(ref_mt, name) is not getattr(our_mt, name)
    kw = dict(lr=1e-3, test_step=1, lr_decay=0.95, model=name, loss="CrossEntropy", n_classes=6, dataset="IEMOCAP")
    torch.manual_seed(111)
    ref = ref_mt.ModelTrainer(device=torch.device("cpu"), **kw)
    torch.manual_seed(111)
    ours = our_mt.ModelTrainer(device=torch.device("cuda"), **kw)
    ours.load_parameters_from = None
    # same initial weights by construction (same RNG order); perturb the ones-initialised tensors identically (well-conditioned)
    from oracle import torch_port as tp
    # sps couples the dialogues of a batch through packed rows (F3) and needs one full-length dialogue per batch
    batches = _loader(5, [[9, 4, 7, 9, 5], [8, 8, 3, 6]])
    tp.perturb_ones(ref.model, _kink_free_perturbation(name, ref.state_dict(), batches))
    ours.load_state_dict({k: v.clone() for k, v in ref.state_dict().items()})
    acc_r, f1_r, _ = ref.eval_network(batches)
    acc_o, f1_o, _ = ours.eval_network(batches)
    assert (acc_r, f1_r) == (acc_o, f1_o), ((acc_r, f1_r), (acc_o, f1_o))
    for t in (ref, ours):
        for m in t.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    before = {k: v.clone() for k, v in ref.state_dict().items()}
    lr_r, loss_r = ref.train_network(1, batches)
    lr_o, loss_o = ours.train_network(1, batches)
    assert lr_r == lr_o and abs(loss_r - loss_o) <= 2e-4 * max(1.0, abs(loss_r)), (loss_r, loss_o)
    sr, so = ref.state_dict(), ours.state_dict()
    # Adam's update is lr * m / (sqrt(v) + eps): an element whose gradient is rounding noise still moves by ~lr in a direction
    # the noise decides, so element-wise equality of the weights is not a meaningful bar.  The bar is on the update vector
    # of the whole model (2 % in L2) plus the trivial per-element bound.
    num = den = 0.0
    worst = 0.0
    contrib = {}
    for k in sr:
        du_r, du_o = (sr[k] - before[k]).double(), (so[k].cpu() - before[k]).double()
        num += float((du_o - du_r).pow(2).sum())
        den += float(du_r.pow(2).sum())
        worst = max(worst, float((du_o - du_r).abs().max()))
        contrib[k] = (float((du_o - du_r).pow(2).sum()), float(du_r.pow(2).sum()))
        assert float((so[k].cpu() - sr[k]).abs().max()) <= 5e-3, k     # both sides move <= ~2.4 lr in two steps
    top = sorted(contrib.items(), key=lambda kv: -kv[1][0])[:4]
    assert den > 0 and (num / den) ** 0.5 <= 2e-2, ((num / den) ** 0.5, top)
    never_used = [n for n, p in ref.model.named_parameters() if p.grad is None]
    assert never_used and all(dict(ours.model.named_parameters())[n].grad is None for n in never_used)
    print(name, "epoch loss", loss_r, loss_o, "relative L2 of the update difference", (num / den) ** 0.5, "max element", worst, "unused tensors", len(never_used))
