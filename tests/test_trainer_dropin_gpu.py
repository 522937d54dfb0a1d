"""Drop-in boundary under the reference's OWN training loop: the unmodified
``model_trainer.py`` (staged byte for byte in oracle/_ref) builds the model by name, owns Adam + StepLR, and runs
``train_network`` / ``eval_network``.  The only integration is that its ``from models.lsthm_<x> import MARN1_<x>`` lines
resolve to the mirror modules.  Checked against the same trainer driving the reference classes on the CPU:

  * ``eval_network``: identical predictions (accuracy, weighted F1) on a synthetic IEMOCAP-shaped loader;
  * ``train_network`` (dropout probabilities set to 0 on both sides so the two runs are comparable): the epoch loss over two
    optimizer steps, and every parameter after the two Adam steps — forward, backward and the optimizer through the
    reference's code path; parameters the reference never uses keep ``grad = None`` and are not decayed (F8).
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


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference neither at /root/reference nor staged in oracle/_ref")
@pytest.mark.parametrize("name", ["MARN1_onlysp", "MARN1_nsps", "MARN1_sps"])
def test_reference_trainer_drives_the_dropin(name, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                                   # eval_network writes res.csv into the cwd (model_trainer.py:158)
    mod = {"MARN1_onlysp": "lsthm_onlysp", "MARN1_nsps": "lsthm_nsps", "MARN1_sps": "lsthm_sps"}[name]
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
    tp.perturb_ones(ref.model, 7)
    ours.load_state_dict({k: v.clone() for k, v in ref.state_dict().items()})
    # sps couples the dialogues of a batch through packed rows (F3) and needs one full-length dialogue per batch
    batches = _loader(5, [[9, 4, 7, 9, 5], [8, 8, 3, 6]])
    acc_r, f1_r, _ = ref.eval_network(batches)
    acc_o, f1_o, _ = ours.eval_network(batches)
    assert (acc_r, f1_r) == (acc_o, f1_o), ((acc_r, f1_r), (acc_o, f1_o))
    for t in (ref, ours):
        for m in t.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    lr_r, loss_r = ref.train_network(1, batches)
    lr_o, loss_o = ours.train_network(1, batches)
    assert lr_r == lr_o and abs(loss_r - loss_o) <= 2e-4 * max(1.0, abs(loss_r)), (loss_r, loss_o)
    sr, so = ref.state_dict(), ours.state_dict()
    worst = 0.0
    for k in sr:
        d = float((so[k].cpu() - sr[k]).abs().max())
        worst = max(worst, d)
        # two Adam steps of lr 1e-3 move a weight by <= 2e-3; parity of the update direction to 5 % of that
        assert d <= 1e-4, (k, d)
    never_used = [n for n, p in ref.model.named_parameters() if p.grad is None]
    assert never_used and all(dict(ours.model.named_parameters())[n].grad is None for n in never_used)
    print(name, "epoch loss", loss_r, loss_o, "max param diff after 2 Adam steps", worst, "unused tensors", len(never_used))
