"""Drop-in for the reference's ``model/HybridRNN_AT.py``: ``MARN()`` -> forward(x[L,B,200]) ->
probabilities [L*B, 7] (time-major), text 100 + audio 100 (HybridRNN_AT.py:40-144)."""
from .mab_net import LSTHM, MabNet  # noqa: F401


class MARN(MabNet):
    def __init__(self):
        super().__init__(d_in=(100, 100), dh=(128, 16), reduce=(16, 128), output_dim=7)
        self.d_l, self.d_a = self._d_in
        self.dh_l, self.dh_a = self._dh
