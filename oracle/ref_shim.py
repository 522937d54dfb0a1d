"""TEST INFRASTRUCTURE — import shim for the *live* reference.

The reference at /root/reference is pure Python/PyTorch but its import paths are
broken as shipped: code says ``models.*`` / ``attention.*`` while the
directories are ``model/`` and ``attention:/``; several files ``import imp`` (gone in
Python 3.12).  This shim repairs those three things *in sys.modules only* — nothing
is copied out of /root/reference and nothing is written into it.

/root/reference does not exist on the GPU box.  ``make -C oracle ref`` (run by ``__graft_entry__.build()`` in the build
container) stages byte-for-byte copies of the hot-path files into the git-ignored ``oracle/_ref/`` — with a SHA256SUMS
manifest — which does travel to the GPU box; the shim uses /root/reference when present and the staged tree otherwise.
Users: ``oracle/make_golden.py``, the tests, and ``bench.py``'s reference arm / ``cpu_baseline`` (the unmodified
reference, timed; never the product path).
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = os.environ.get("LSTHM_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "model")) and os.path.isdir(os.path.join(_STAGED, "model")):
    REF_ROOT = _STAGED
REF_KIND = "staged" if REF_ROOT == _STAGED else "tree"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def _attention_dir() -> str:
    a = os.path.join(REF_ROOT, "attention:")
    return a if os.path.isdir(a) else os.path.join(REF_ROOT, "attention_")


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
        _alias_package("attention", _attention_dir())
    ns = types.SimpleNamespace()
    from models.HybridRNN_ATV import MARN as MARN_ATV  # model/HybridRNN_ATV.py:40
    from models.HybridRNN_AT import MARN as MARN_AT    # model/HybridRNN_AT.py:40
    from models.lsthm_sps import MARN1_sps              # model/lsthm_sps.py:298
    from models.lsthm_onlysp import MARN1_onlysp        # model/lsthm_onlysp.py:213 (train.py default model)
    from models.lsthm_nsps import MARN1_nsps            # model/lsthm_nsps.py:283
    from models.lsthm_no_en import MARN1_no_en          # model/lsthm_no_en.py:283 (nsps without the text encoder)
    from models.encoder import EncoderLayer             # model/encoder.py:116
    from models.DialogueRNN import BiModel              # model/DialogueRNN.py:201 (config 4 baseline)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_lsthm_ref_loss", os.path.join(REF_ROOT, "loss.py"))
    ref_loss = importlib.util.module_from_spec(spec)    # loss.py:6 (loaded by path: our package also has a loss.py)
    spec.loader.exec_module(ref_loss)
    ns.MARN_ATV, ns.MARN_AT, ns.MARN1_sps, ns.MARN1_onlysp = MARN_ATV, MARN_AT, MARN1_sps, MARN1_onlysp
    ns.MARN1_nsps, ns.MARN1_no_en, ns.BiModel = MARN1_nsps, MARN1_no_en, BiModel
    ns.EncoderLayer, ns.MaskedLoss = EncoderLayer, ref_loss.MaskedLoss
    ns.root, ns.kind = REF_ROOT, REF_KIND
    return ns


def load_trainer(overrides=None, name="_lsthm_ref_model_trainer"):
    """The reference's UNMODIFIED ``model_trainer.py`` (ModelTrainer: model construction by name, Adam + StepLR,
    train_network / eval_network, model_trainer.py:29-187) as a module object.  ``overrides`` maps ``models.<file>`` module
    names to replacement modules that are visible while the trainer's own ``from models.<file> import <Class>`` lines run —
    that is the whole integration of INTEGRATION.md §2a: the trainer then builds and trains the drop-in classes.
    ``librosa`` / ``soundfile`` (imported, never used by the trainer) are stubbed when absent."""
    import importlib.util
    load_reference()
    for m in ("librosa", "soundfile"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)
    spec = importlib.util.spec_from_file_location("_lsthm_ref_loss", os.path.join(REF_ROOT, "loss.py"))
    ref_loss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_loss)
    saved = {k: sys.modules.get(k) for k in list(overrides or {}) + ["loss"]}
    try:
        sys.modules["loss"] = ref_loss                        # `from loss import MaskedLoss, InfoNCE` (model_trainer.py:13)
        for k, v in (overrides or {}).items():
            sys.modules[k] = v
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, "model_trainer.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


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
