"""Drop-in for the reference's ``model/HybridRNN_ATV.py``: ``MARN()`` -> forward(x[L,B,712]) ->
probabilities [L*B, 6] (time-major), text 100 + audio 100 + visual 512 (HybridRNN_ATV.py:40-155)."""
from .mab_net import LSTHM, MabNet  # noqa: F401


class MARN(MabNet):
    def __init__(self):
        super().__init__(d_in=(100, 100, 512), dh=(128, 16, 64), reduce=(16, 128, 100), output_dim=6)
        self.d_l, self.d_a, self.d_v = self._d_in
        self.dh_l, self.dh_a, self.dh_v = self._dh
