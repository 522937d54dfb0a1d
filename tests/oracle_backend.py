"""Test-only stand-in for the CUDA library: routes recurrence.py's ABI calls to the fp64 plain-C
oracle so the HOST logic (autograd plumbing, hoisted weight-gradient products, module wiring) can
be checked on a machine without a GPU.  Installed by monkeypatching inside a test; the product
never imports this file (it lives under tests/)."""
from importlib import import_module

import numpy as np
import torch

from oracle import cpu as ocpu

MODS = ("l", "a", "v")


class OracleBackend:
    def __init__(self):
        self.weights = None

    # --- signatures mirror _lib.py ---
    def make_weights(self, U, V, Watt, batt, Wr, br, Wf1, bf1, Wf2, bf2):
        p = {}
        for i in range(len(U)):
            m = MODS[i]
            p[f"lsthm_{m}.U.weight"], p[f"lsthm_{m}.V.weight"] = U[i], V[i]
            p[f"reduce_dim_nn_{m}.0.weight"], p[f"reduce_dim_nn_{m}.0.bias"] = Wr[i], br[i]
        p["att.0.weight"], p["att.0.bias"] = Watt, batt
        p["fc.0.weight"], p["fc.0.bias"], p["fc.3.weight"], p["fc.3.bias"] = Wf1, bf1, Wf2, bf2
        self.weights = {k: v.detach().double().numpy() for k, v in p.items()}
        return self.weights

    def mab_pack_bytes(self, d):
        return 128

    def mab_workspace_bytes(self, d):
        return 128

    def mab_pack(self, d, w, packed):
        pass

    @staticmethod
    def _dims(d):
        return tuple(d.dh[i] for i in range(d.n_mod)), tuple(d.rd[i] for i in range(d.n_mod))

    def mab_alloc_stash(self, d, device):
        """The stash is private to the fwd/bwd pair: this stand-in keeps the oracle's own tensors (c, gates, softmax)."""
        dh, _ = self._dims(d)
        D = sum(dh)
        z = lambda w: torch.zeros(d.T * d.N * w)
        return dict(sCp=z(D), sG=z(4 * D), sE=z(4 * D), sMS=z(8), sP=z(4 * d.map_h))

    def mab_fwd(self, d, packed, gx, drop_mask, hz, u, sC, sCp, sG, sE, sMS, sP, workspace):
        dh, rd = self._dims(d)
        f = ocpu.mab_forward(self.weights, gx.double().numpy(), dh, rd,
                             None if drop_mask is None else drop_mask.double().numpy(), d.map_h)
        hz.copy_(torch.from_numpy(f["hz"]))            # the library writes only the h half; the host overwrites the z half
        u.copy_(torch.from_numpy(f["UH"]).reshape(u.shape))
        if sC is not None:
            sC.copy_(torch.from_numpy(f["C"]).reshape(sC.shape))
            sCp.copy_(torch.from_numpy(f["C"]).reshape(-1))
            sG.copy_(torch.from_numpy(f["G"]).reshape(-1))
            sE.copy_(torch.from_numpy(f["A"]).reshape(-1))       # (this stand-in stashes the softmax weights themselves)

    def mab_bwd(self, d, packed, dhz, duz, drop_mask, sCp, sG, sE, sMS, sP, u, dgx, de, dup, att, workspace):
        dh, rd = self._dims(d)
        T, N = d.T, d.N
        D = sum(dh)
        sC, sA = sCp.view(T, N, D), sE.view(T, N, 4, D)
        # hz / R feed only the oracle's own weight-gradient accumulation, which is not used here; the oracle takes the
        # head's dL/dz directly (dhz carries it), so `duz` (= dz_head Wf2, what the CUDA kernel consumes) is not needed
        f = dict(hz=np.zeros((T, N, 2 * D)), C=sC.double().numpy(), G=sG.view(T, N, 4 * D).double().numpy(),
                 A=np.ascontiguousarray(sA.double().numpy()),
                 R=np.zeros((T, N, sum(rd))), UH=u.double().numpy())
        adj, _ = ocpu.mab_backward(self.weights, dhz.double().numpy(), f, dh, rd,
                                   None if drop_mask is None else drop_mask.double().numpy(), d.map_h)
        for t, k in ((dgx, "dgx"), (de, "de"), (dup, "dup")):
            t.copy_(torch.from_numpy(adj[k]).reshape(t.shape))
        if att is not None:   # attended = a * c regrouped per modality, head-major (include/lsthm_b200.h: lsthm_mab_bwd)
            a4 = sA * sC.reshape(T, N, 1, D)
            o = 0
            for h in dh:
                att[:, :, 4 * o:4 * o + 4 * h] = a4[:, :, :, o:o + h].reshape(T, N, 4 * h)
                o += h


def install(monkeypatch):
    import lsthm_b200
    lib = import_module(lsthm_b200.__name__ + "._lib")
    rec = import_module(lsthm_b200.__name__ + ".recurrence")
    net = import_module(lsthm_b200.__name__ + ".mab_net")
    be = OracleBackend()
    for name in ("make_weights", "mab_pack_bytes", "mab_workspace_bytes", "mab_pack", "mab_alloc_stash", "mab_fwd", "mab_bwd"):
        monkeypatch.setattr(lib, name, getattr(be, name))
    monkeypatch.setattr(rec, "_workspace", lambda desc, device: torch.zeros(128, dtype=torch.uint8))

    # the public entry point refuses CPU tensors; tests go through the autograd.Function directly
    def cpu_recurrence(gx, mask, dh, rd, map_h, weights, rows=0, prepacked=None):
        assert prepacked is None                 # side-stream prepacking is a CUDA-only schedule
        return rec.MabRecurrenceFn.apply(gx, mask, (tuple(dh), tuple(rd), int(map_h), int(rows)), *weights)

    monkeypatch.setattr(net, "mab_recurrence", cpu_recurrence)
    return be


def install_plain():
    """Same as install() but without pytest's monkeypatch (for spawned worker processes)."""
    class _MP:
        @staticmethod
        def setattr(obj, name, value):
            setattr(obj, name, value)
    return install(_MP)
