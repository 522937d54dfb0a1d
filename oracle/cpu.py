"""TEST INFRASTRUCTURE — ctypes front end of oracle/mab_oracle.c (fp64 plain-C restatement of
the AT/ATV recurrence forward + hand-derived BPTT).  Checker only: imported by tests/,
``__graft_entry__.smoke()`` and nothing on the product path."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmab_oracle.so")
MAXM = 3


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mab_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _SO


class _Dims(C.Structure):
    _fields_ = [("T", C.c_int), ("N", C.c_int), ("n_mod", C.c_int), ("n_att", C.c_int),
                ("map_h", C.c_int), ("dh", C.c_int * MAXM), ("rd", C.c_int * MAXM)]


_P = C.POINTER(C.c_double)


class _Weights(C.Structure):
    _fields_ = [("U", _P * MAXM), ("V", _P * MAXM), ("Watt", _P), ("batt", _P),
                ("Wr", _P * MAXM), ("br", _P * MAXM), ("Wf1", _P), ("bf1", _P), ("Wf2", _P), ("bf2", _P)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        assert _lib.mab_oracle_real_bytes() == 8
    return _lib


def _ptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P)


MODS = ("l", "a", "v")


def pack_weights(params: dict, n_mod: int, keep: list):
    """params: name -> array-like (state_dict names of the reference).  Returns (_Weights, dict of f64 arrays)."""
    w = _Weights()
    arrs = {}

    def get(name):
        a = np.ascontiguousarray(np.asarray(params[name], dtype=np.float64))
        arrs[name] = a
        keep.append(a)
        return _ptr(a)

    for k in range(n_mod):
        m = MODS[k]
        w.U[k] = get(f"lsthm_{m}.U.weight")
        w.V[k] = get(f"lsthm_{m}.V.weight")
        w.Wr[k] = get(f"reduce_dim_nn_{m}.0.weight")
        w.br[k] = get(f"reduce_dim_nn_{m}.0.bias")
    w.Watt, w.batt = get("att.0.weight"), get("att.0.bias")
    w.Wf1, w.bf1 = get("fc.0.weight"), get("fc.0.bias")
    w.Wf2, w.bf2 = get("fc.3.weight"), get("fc.3.bias")
    return w, arrs


def make_dims(T, N, dh, rd, map_h=64, n_att=4):
    d = _Dims()
    d.T, d.N, d.n_mod, d.n_att, d.map_h = T, N, len(dh), n_att, map_h
    for i, (a, b) in enumerate(zip(dh, rd)):
        d.dh[i], d.rd[i] = a, b
    return d


def mab_forward(params: dict, gx: np.ndarray, dh, rd, drop_mask=None, map_h=64, n_att=4):
    """gx [T,N,4D] f64.  Returns dict(hz, C, G, A, R, UH)."""
    T, N, G = gx.shape
    D, RD = sum(dh), sum(rd)
    assert G == 4 * D
    keep = []
    w, _ = pack_weights(params, len(dh), keep)
    d = make_dims(T, N, dh, rd, map_h, n_att)
    gx = np.ascontiguousarray(gx, dtype=np.float64)
    out = dict(hz=np.zeros((T, N, 2 * D)), C=np.zeros((T, N, D)), G=np.zeros((T, N, 4 * D)),
               A=np.zeros((T, N, n_att, D)), R=np.zeros((T, N, RD)), UH=np.zeros((T, N, map_h)))
    dm = None
    if drop_mask is not None:
        dm = np.ascontiguousarray(drop_mask, dtype=np.float64)
    lib().mab_oracle_fwd(C.byref(d), C.byref(w), _ptr(gx), _ptr(dm) if dm is not None else None,
                         _ptr(out["hz"]), _ptr(out["C"]), _ptr(out["G"]), _ptr(out["A"]),
                         _ptr(out["R"]), _ptr(out["UH"]))
    return out


class _WGrads(C.Structure):
    _fields_ = [("U", _P * MAXM), ("V", _P * MAXM), ("Watt", _P), ("batt", _P),
                ("Wr", _P * MAXM), ("br", _P * MAXM), ("Wf1", _P), ("bf1", _P), ("Wf2", _P), ("bf2", _P)]


def mab_backward(params: dict, dhz: np.ndarray, fwd: dict, dh, rd, drop_mask=None, map_h=64, n_att=4):
    """Returns (adjoint dict: dgx, de, dr, dup, dzt ; weight-grad dict keyed by state_dict name)."""
    T, N, _ = dhz.shape
    D, RD = sum(dh), sum(rd)
    keep = []
    w, arrs = pack_weights(params, len(dh), keep)
    d = make_dims(T, N, dh, rd, map_h, n_att)
    adj = dict(dgx=np.zeros((T, N, 4 * D)), de=np.zeros((T, N, n_att, D)), dr=np.zeros((T, N, RD)),
               dup=np.zeros((T, N, map_h)), dzt=np.zeros((T, N, D)))
    gw = _WGrads()
    grads = {name: np.zeros_like(a) for name, a in arrs.items()}
    for k in range(len(dh)):
        m = MODS[k]
        gw.U[k], gw.V[k] = _ptr(grads[f"lsthm_{m}.U.weight"]), _ptr(grads[f"lsthm_{m}.V.weight"])
        gw.Wr[k], gw.br[k] = _ptr(grads[f"reduce_dim_nn_{m}.0.weight"]), _ptr(grads[f"reduce_dim_nn_{m}.0.bias"])
    gw.Watt, gw.batt = _ptr(grads["att.0.weight"]), _ptr(grads["att.0.bias"])
    gw.Wf1, gw.bf1 = _ptr(grads["fc.0.weight"]), _ptr(grads["fc.0.bias"])
    gw.Wf2, gw.bf2 = _ptr(grads["fc.3.weight"]), _ptr(grads["fc.3.bias"])
    dm = np.ascontiguousarray(drop_mask, dtype=np.float64) if drop_mask is not None else None
    dhz = np.ascontiguousarray(dhz, dtype=np.float64)
    lib().mab_oracle_bwd(C.byref(d), C.byref(w), _ptr(dhz), _ptr(dm) if dm is not None else None,
                         _ptr(fwd["hz"]), _ptr(fwd["C"]), _ptr(fwd["G"]), _ptr(fwd["A"]), _ptr(fwd["R"]),
                         _ptr(fwd["UH"]), _ptr(adj["dgx"]), _ptr(adj["de"]), _ptr(adj["dr"]),
                         _ptr(adj["dup"]), _ptr(adj["dzt"]), C.byref(gw))
    return adj, grads
