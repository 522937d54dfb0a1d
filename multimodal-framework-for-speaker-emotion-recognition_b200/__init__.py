"""B200-native LSTHM hybrid recurrence (hot path of
MallVilliers/Multimodal-Framework-for-speaker-emotion-recognition) behind the reference's own
module API.  Import by name with importlib (the directory name has hyphens) or through the
``lsthm_b200`` alias module at the repo root."""
from . import _lib  # noqa: F401
from .recurrence import mab_recurrence, launch_counter  # noqa: F401
from . import HybridRNN_AT, HybridRNN_ATV, lsthm_sps, lsthm_onlysp, lsthm_nsps, lsthm_no_en  # noqa: F401
from .loss import MaskedLoss  # noqa: F401

__all__ = ["HybridRNN_AT", "HybridRNN_ATV", "lsthm_sps", "lsthm_onlysp", "lsthm_nsps", "lsthm_no_en", "MaskedLoss", "mab_recurrence", "launch_counter"]
