"""TEST INFRASTRUCTURE — import shim for the *live* reference (container only).

The reference at /root/reference is pure Python/PyTorch but its import paths are
broken as shipped: code says ``models.*`` / ``attention.*`` while the
directories are ``model/`` and ``attention:/``; several files ``import imp`` (gone in
Python 3.12).  This shim repairs those three things *in sys.modules only* — nothing
is copied out of /root/reference and nothing is written into it.

/root/reference does not exist on the GPU box: only ``oracle/make_golden.py`` and the
``-m "not gpu"`` tests that are explicitly skipped when the directory is absent may
call :func:`load_reference`.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("LSTHM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def _alias_package(name: str, path: str) -> None:
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod


def load_reference():
    """Make ``models.HybridRNN_ATV`` & co importable; returns a namespace of classes."""
    if not reference_available():
        raise RuntimeError(f"reference tree not present at {REF_ROOT}")
    sys.modules.setdefault("imp", types.ModuleType("imp"))
    if "models" not in sys.modules:
        _alias_package("models", os.path.join(REF_ROOT, "model"))
    if "attention" not in sys.modules:
        _alias_package("attention", os.path.join(REF_ROOT, "attention:"))
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)  # for loss.py
    ns = types.SimpleNamespace()
    from models.HybridRNN_ATV import MARN as MARN_ATV  # model/HybridRNN_ATV.py:40
    from models.HybridRNN_AT import MARN as MARN_AT    # model/HybridRNN_AT.py:40
    from models.lsthm_sps import MARN1_sps              # model/lsthm_sps.py:298
    from models.lsthm_onlysp import MARN1_onlysp        # model/lsthm_onlysp.py:213 (train.py default model)
    from models.encoder import EncoderLayer             # model/encoder.py:116
    import loss as ref_loss                             # loss.py:6
    ns.MARN_ATV, ns.MARN_AT, ns.MARN1_sps, ns.MARN1_onlysp = MARN_ATV, MARN_AT, MARN1_sps, MARN1_onlysp
    ns.EncoderLayer, ns.MaskedLoss = EncoderLayer, ref_loss.MaskedLoss
    return ns


def attach_tape(model, tape):
    """Replace every nn.Dropout inside a *reference* model by a tape-driven one whose site
    name is the module path (e.g. ``fc.2``), so reference and oracle/CUDA path consume the
    same masks in train mode."""
    import torch.nn as nn

    class _TapeDropout(nn.Module):
        def __init__(self, site, p):
            super().__init__()
            self.site, self.p = site, p

        def forward(self, x):
            if not self.training or self.p == 0.0:
                return x
            return x * tape.mask(self.site, x.shape, self.p, x.dtype)

    for path, mod in list(model.named_modules()):
        for child_name, child in list(mod.named_children()):
            if isinstance(child, nn.Dropout):
                site = f"{path}.{child_name}" if path else child_name
                setattr(mod, child_name, _TapeDropout(site, child.p))
    return model
