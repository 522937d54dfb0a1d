"""ctypes binding of liblsthm_b200.so (the C ABI in include/lsthm_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
caller gets a RuntimeError.  PyTorch is used only to own device memory and to name the
current CUDA stream; no torch type crosses the ABI.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# LSTHM_B200_SO lets profiling scripts load an experimental build of the same ABI (never a different backend)
SO_PATH = os.environ.get("LSTHM_B200_SO") or os.path.join(_PKG, "liblsthm_b200.so")
ABI_VERSION = 7
MAX_MOD = 3

_f32p = C.POINTER(C.c_float)


class MabDesc(C.Structure):
    _fields_ = [("T", C.c_int32), ("N", C.c_int32), ("n_mod", C.c_int32), ("n_att", C.c_int32),
                ("map_h", C.c_int32), ("dh", C.c_int32 * MAX_MOD), ("rd", C.c_int32 * MAX_MOD),
                ("rows_per_cta", C.c_int32)]


class MabWeights(C.Structure):
    _fields_ = [("U", C.c_void_p * MAX_MOD), ("V", C.c_void_p * MAX_MOD), ("Watt", C.c_void_p),
                ("batt", C.c_void_p), ("Wr", C.c_void_p * MAX_MOD), ("br", C.c_void_p * MAX_MOD),
                ("Wf1", C.c_void_p), ("bf1", C.c_void_p), ("Wf2", C.c_void_p), ("bf2", C.c_void_p)]


class SpsDesc(C.Structure):
    _fields_ = [("T", C.c_int32), ("N", C.c_int32), ("rows_per_cta", C.c_int32), ("att_p", C.c_float),
                ("att_seed", C.c_uint64)]


class SpsWeights(C.Structure):
    _fields_ = [("U", C.c_void_p * 2), ("V", C.c_void_p * 2), ("S", C.c_void_p * 2), ("Wih", C.c_void_p * 2),
                ("Whh", C.c_void_p * 2), ("bq", C.c_void_p * 2), ("Wq", C.c_void_p), ("Wk", C.c_void_p)]


class SpsMasks(C.Structure):
    _fields_ = [("mq", C.c_void_p * 2), ("ml", C.c_void_p), ("ma", C.c_void_p), ("att_mask", C.c_void_p)]


class GspDesc(C.Structure):
    _fields_ = [("T", C.c_int32), ("N", C.c_int32), ("rows_per_cta", C.c_int32), ("listener", C.c_int32),
                ("att_p", C.c_float), ("att_seed", C.c_uint64)]


class GspWeights(C.Structure):
    _fields_ = [("U", C.c_void_p * 2), ("V", C.c_void_p * 2), ("S", C.c_void_p * 2), ("Whh", C.c_void_p),
                ("bhh", C.c_void_p), ("Wq", C.c_void_p), ("Wk", C.c_void_p)]


class GspMasks(C.Structure):
    _fields_ = [("ms", C.c_void_p), ("ml", C.c_void_p), ("ma", C.c_void_p), ("att_mask", C.c_void_p)]


class AttnDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("d_head", C.c_int32), ("ldq", C.c_int32),
                ("ldk", C.c_int32), ("ldv", C.c_int32), ("ldo", C.c_int32), ("scale", C.c_float), ("p_drop", C.c_float),
                ("seed", C.c_uint64), ("row_stride_b", C.c_int64), ("row_stride_i", C.c_int64),
                ("precision", C.c_int32), ("reserved", C.c_int32)]


class XAttnDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("D", C.c_int32), ("ldq", C.c_int32), ("ldk", C.c_int32),
                ("ldv", C.c_int32), ("ldo", C.c_int32), ("scale", C.c_float), ("p_drop", C.c_float), ("seed", C.c_uint64),
                ("row_stride_b", C.c_int64), ("row_stride_i", C.c_int64)]


class DlnDesc(C.Structure):
    _fields_ = [("R", C.c_int64), ("d", C.c_int32), ("eps", C.c_float), ("p_drop", C.c_float), ("seed", C.c_uint64)]


def build(verbose: bool = False, jobs: int = 8) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_PKG, "csrc"), f"-j{jobs}", "all"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode:
        print(res.stdout)
    if res.returncode:
        raise RuntimeError("building liblsthm_b200.so failed (see output above)")
    return SO_PATH


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: the CUDA extension is not built and there is no CPU fallback. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc).")
    L = C.CDLL(SO_PATH)
    L.lsthm_abi_version.restype = C.c_int
    L.lsthm_last_error.restype = C.c_char_p
    L.lsthm_mab_pack_bytes.restype = C.c_size_t
    L.lsthm_mab_pack_bytes.argtypes = [C.POINTER(MabDesc)]
    L.lsthm_mab_workspace_bytes.restype = C.c_size_t
    L.lsthm_mab_workspace_bytes.argtypes = [C.POINTER(MabDesc)]
    L.lsthm_mab_pack.restype = C.c_int
    L.lsthm_mab_pack.argtypes = [C.POINTER(MabDesc), C.POINTER(MabWeights), C.c_void_p, C.c_void_p]
    L.lsthm_mab_fwd.restype = C.c_int
    L.lsthm_mab_fwd.argtypes = [C.POINTER(MabDesc)] + [C.c_void_p] * 13
    L.lsthm_mab_bwd.restype = C.c_int
    L.lsthm_mab_bwd.argtypes = [C.POINTER(MabDesc)] + [C.c_void_p] * 16
    L.lsthm_mab_launch_info.restype = C.c_int
    L.lsthm_mab_launch_info.argtypes = [C.POINTER(MabDesc)] + [C.POINTER(C.c_int32)] * 7
    L.lsthm_mab_plan_info.restype = C.c_int
    L.lsthm_mab_plan_info.argtypes = [C.POINTER(MabDesc), C.POINTER(C.c_int32), C.c_int32]
    L.lsthm_mab_set_trace.restype = C.c_int
    L.lsthm_mab_set_trace.argtypes = [C.c_void_p]
    L.lsthm_sps_packed_floats.restype = C.c_size_t
    L.lsthm_sps_packed_floats.argtypes = []
    L.lsthm_sps_workspace_floats.restype = C.c_size_t
    L.lsthm_sps_workspace_floats.argtypes = [C.POINTER(SpsDesc)]
    L.lsthm_sps_pack.restype = C.c_int
    L.lsthm_sps_pack.argtypes = [C.POINTER(SpsWeights), C.c_void_p, C.c_void_p]
    L.lsthm_sps_fwd.restype = C.c_int
    L.lsthm_sps_fwd.argtypes = [C.POINTER(SpsDesc), C.POINTER(SpsWeights)] + [C.c_void_p] * 5 + [C.POINTER(SpsMasks)] + [C.c_void_p] * 10
    L.lsthm_sps_bwd.restype = C.c_int
    L.lsthm_sps_bwd.argtypes = [C.POINTER(SpsDesc), C.POINTER(SpsWeights)] + [C.c_void_p] * 4 + [C.POINTER(SpsMasks)] + [C.c_void_p] * 10
    L.lsthm_sps_launch_info.restype = C.c_int
    L.lsthm_sps_launch_info.argtypes = [C.POINTER(SpsDesc)] + [C.POINTER(C.c_int32)] * 5
    L.lsthm_gsp_packed_floats.restype = C.c_size_t
    L.lsthm_gsp_packed_floats.argtypes = []
    L.lsthm_gsp_pack.restype = C.c_int
    L.lsthm_gsp_pack.argtypes = [C.POINTER(GspWeights), C.c_void_p, C.c_void_p]
    L.lsthm_gsp_fwd.restype = C.c_int
    L.lsthm_gsp_fwd.argtypes = [C.POINTER(GspDesc), C.POINTER(GspWeights)] + [C.c_void_p] * 4 + [C.POINTER(GspMasks)] + [C.c_void_p] * 6
    L.lsthm_gsp_bwd.restype = C.c_int
    L.lsthm_gsp_bwd.argtypes = [C.POINTER(GspDesc), C.POINTER(GspWeights), C.c_void_p, C.POINTER(GspMasks)] + [C.c_void_p] * 10
    L.lsthm_gsp_launch_info.restype = C.c_int
    L.lsthm_gsp_launch_info.argtypes = [C.POINTER(GspDesc)] + [C.POINTER(C.c_int32)] * 5
    L.lsthm_gemm3_workspace_floats.restype = C.c_size_t
    L.lsthm_gemm3_workspace_floats.argtypes = [C.c_int32] * 4
    L.lsthm_gemm3.restype = C.c_int
    L.lsthm_gemm3.argtypes = [C.c_int32] * 4 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                                C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    L.lsthm_attn_fwd.restype = C.c_int
    L.lsthm_attn_fwd.argtypes = [C.POINTER(AttnDesc)] + [C.c_void_p] * 6
    L.lsthm_attn_bwd.restype = C.c_int
    L.lsthm_attn_bwd.argtypes = [C.POINTER(AttnDesc)] + [C.c_void_p] * 10
    L.lsthm_xattn_fwd.restype = C.c_int
    L.lsthm_xattn_fwd.argtypes = [C.POINTER(XAttnDesc)] + [C.c_void_p] * 6
    L.lsthm_xattn_bwd.restype = C.c_int
    L.lsthm_xattn_bwd.argtypes = [C.POINTER(XAttnDesc)] + [C.c_void_p] * 8
    L.lsthm_reverse_seq.restype = C.c_int
    L.lsthm_reverse_seq.argtypes = [C.c_int32] * 3 + [C.c_void_p] * 4
    L.lsthm_masked_loss_workspace_floats.restype = C.c_size_t
    L.lsthm_masked_loss_workspace_floats.argtypes = [C.c_int64]
    L.lsthm_masked_loss_fwd.restype = C.c_int
    L.lsthm_masked_loss_fwd.argtypes = [C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 6
    L.lsthm_masked_loss_bwd.restype = C.c_int
    L.lsthm_masked_loss_bwd.argtypes = [C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 7
    L.lsthm_dln_workspace_floats.restype = C.c_size_t
    L.lsthm_dln_workspace_floats.argtypes = [C.c_int32]
    L.lsthm_dln_fwd.restype = C.c_int
    L.lsthm_dln_fwd.argtypes = [C.POINTER(DlnDesc), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    L.lsthm_dln_bwd.restype = C.c_int
    L.lsthm_dln_bwd.argtypes = [C.POINTER(DlnDesc), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.lsthm_gemm3w_pack_bytes.restype = C.c_size_t
    L.lsthm_gemm3w_pack_bytes.argtypes = [C.c_int32, C.c_int32]
    L.lsthm_gemm3w.restype = C.c_int
    L.lsthm_gemm3w.argtypes = [C.c_int32] * 4 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                               C.c_void_p, C.c_size_t, C.c_void_p]
    L.lsthm_colsum_workspace_floats.restype = C.c_size_t
    L.lsthm_colsum_workspace_floats.argtypes = [C.c_int64, C.c_int32]
    L.lsthm_colsum.restype = C.c_int
    L.lsthm_colsum.argtypes = [C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.lsthm_assemble_input.restype = C.c_int
    L.lsthm_assemble_input.argtypes = [C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 7
    L.lsthm_adam_step.restype = C.c_int
    L.lsthm_adam_step.argtypes = [C.c_void_p] * 4 + [C.c_size_t] + [C.c_float] * 5 + [C.c_int32, C.c_void_p]
    if L.lsthm_abi_version() != ABI_VERSION:
        raise RuntimeError(f"liblsthm_b200.so ABI {L.lsthm_abi_version()} != expected {ABI_VERSION}")
    _lib = L
    return L


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib().lsthm_last_error().decode()}")


def _dev_ptr(t: Optional[torch.Tensor], name: str) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    if t.data_ptr() % 16:
        raise RuntimeError(f"{name}: storage must be 16-byte aligned")
    _on_current_device(t, name)
    return t.data_ptr()


def _on_current_device(t: torch.Tensor, name: str) -> None:
    """The library launches on the calling thread's current device and stream: a tensor of another GPU would be
    dereferenced on the wrong device.  Fail loudly instead (wrap the call in ``torch.cuda.device(t.device)``)."""
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"{name}: tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()} "
                           "(call under torch.cuda.device(tensor.device))")


def _stream() -> int:
    """cudaStream_t of the current stream of the current device.  Asked ~110 times per training step: the raw-handle query is
    ~0.3 us, ``torch.cuda.current_stream().cuda_stream`` ~15 us (1.6 ms of host time per ATV step, profiles/dev/host_profile.py)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def make_desc(T: int, N: int, dh: Sequence[int], rd: Sequence[int], map_h: int = 64, n_att: int = 4,
              rows_per_cta: int = 0) -> MabDesc:
    d = MabDesc()
    d.T, d.N, d.n_mod, d.n_att, d.map_h, d.rows_per_cta = T, N, len(dh), n_att, map_h, rows_per_cta
    for i, (a, b) in enumerate(zip(dh, rd)):
        d.dh[i], d.rd[i] = a, b
    return d


def make_weights(U, V, Watt, batt, Wr, br, Wf1, bf1, Wf2, bf2) -> MabWeights:
    w = MabWeights()
    for i in range(len(U)):
        w.U[i], w.V[i] = _dev_ptr(U[i], f"U[{i}]"), _dev_ptr(V[i], f"V[{i}]")
        w.Wr[i], w.br[i] = _dev_ptr(Wr[i], f"Wr[{i}]"), _dev_ptr(br[i], f"br[{i}]")
    w.Watt, w.batt = _dev_ptr(Watt, "Watt"), _dev_ptr(batt, "batt")
    w.Wf1, w.bf1 = _dev_ptr(Wf1, "Wf1"), _dev_ptr(bf1, "bf1")
    w.Wf2, w.bf2 = _dev_ptr(Wf2, "Wf2"), _dev_ptr(bf2, "bf2")
    return w


# ------------------------------------------------------------------------------------------------
# AT / ATV recurrence: weight-stationary tensor-core kernels (lsthm_mab_*)
# ------------------------------------------------------------------------------------------------
def _byte_ptr(t: torch.Tensor, name: str) -> int:
    if not (t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous() and t.data_ptr() % 128 == 0):
        raise RuntimeError(f"{name}: expected a contiguous, 128-byte aligned CUDA uint8 buffer")
    return t.data_ptr()


def mab_pack_bytes(d: MabDesc) -> int:
    n = lib().lsthm_mab_pack_bytes(C.byref(d))
    if n == 0:
        raise RuntimeError(f"unsupported recurrence dims: {lib().lsthm_last_error().decode()}")
    return n


def mab_workspace_bytes(d: MabDesc) -> int:
    n = lib().lsthm_mab_workspace_bytes(C.byref(d))
    if n == 0:
        raise RuntimeError(f"unsupported recurrence dims: {lib().lsthm_last_error().decode()}")
    return n


def mab_pack(d: MabDesc, w: MabWeights, packed: torch.Tensor) -> None:
    _check(lib().lsthm_mab_pack(C.byref(d), C.byref(w), _byte_ptr(packed, "packed"), _stream()), "lsthm_mab_pack")


def mab_fwd(d: MabDesc, packed, gx, drop_mask, hz, u, sC, sCp, sG, sE, sMS, sP, workspace) -> None:
    """Writes the h half of hz[T,N,2D] and u[T,N,map_h]; the caller forms z = u Wf2^T + bf2 (include/lsthm_b200.h)."""
    _check(lib().lsthm_mab_fwd(C.byref(d), _byte_ptr(packed, "packed"), _dev_ptr(gx, "gx"), _dev_ptr(drop_mask, "drop_mask"),
                                _dev_ptr(hz, "hz"), _dev_ptr(u, "u"), _dev_ptr(sC, "sC"), _dev_ptr(sCp, "sCp"), _dev_ptr(sG, "sG"), _dev_ptr(sE, "sE"),
                                _dev_ptr(sMS, "sMS"), _dev_ptr(sP, "sP"), _byte_ptr(workspace, "workspace"), _stream()),
           "lsthm_mab_fwd")


def mab_bwd(d: MabDesc, packed, dhz, duz, drop_mask, sCp, sG, sE, sMS, sP, u, dgx, de, dup, att, workspace) -> None:
    _check(lib().lsthm_mab_bwd(C.byref(d), _byte_ptr(packed, "packed"), _dev_ptr(dhz, "dhz"), _dev_ptr(duz, "duz"),
                                _dev_ptr(drop_mask, "drop_mask"), _dev_ptr(sCp, "sCp"), _dev_ptr(sG, "sG"), _dev_ptr(sE, "sE"),
                                _dev_ptr(sMS, "sMS"), _dev_ptr(sP, "sP"), _dev_ptr(u, "u"), _dev_ptr(dgx, "dgx"), _dev_ptr(de, "de"),
                                _dev_ptr(dup, "dup"), _dev_ptr(att, "att"), _byte_ptr(workspace, "workspace"), _stream()),
           "lsthm_mab_bwd")


def mab_set_trace(buf: Optional[torch.Tensor]) -> None:
    _check(lib().lsthm_mab_set_trace(None if buf is None else buf.data_ptr()), "lsthm_mab_set_trace")


def mab_launch_info(d: MabDesc) -> dict:
    v = [C.c_int32() for _ in range(7)]
    _check(lib().lsthm_mab_launch_info(C.byref(d), *[C.byref(x) for x in v]), "lsthm_mab_launch_info")
    return dict(zip(("grid", "block", "group", "dialogues_per_group", "smem_fwd", "smem_bwd", "padded_rows"), (x.value for x in v)))


def mab_alloc_stash(d: MabDesc, device) -> dict:
    """The private (piece-major) stash tensors of a forward/backward pair, sized for the current device's plan."""
    info = mab_launch_info(d)
    D = sum(d.dh[i] for i in range(d.n_mod))
    rows = d.T * info["padded_rows"]
    new = lambda w: torch.empty(rows * w, device=device, dtype=torch.float32)
    return dict(sCp=new(D), sG=new(4 * D), sE=new(4 * D), sMS=new(8), sP=new(4 * d.map_h))


def mab_unblock(x: torch.Tensor, d: MabDesc, width: int) -> torch.Tensor:
    """Private piece-major stash tensor -> row-major [T, N, width] (tests and debugging only)."""
    info = mab_launch_info(d)
    DG = info["dialogues_per_group"]
    Mr = (DG + 7) // 8 * 8
    nb = info["padded_rows"] // Mr
    if width == 8:          # sMS: [T][block][4 heads][row][2]
        y = x.view(d.T, nb, 4, Mr, 2).permute(0, 1, 3, 2, 4).reshape(d.T, nb, Mr, width)[:, :, :DG]
    else:
        y = x.view(d.T, nb, width // 4, Mr, 4).permute(0, 1, 3, 2, 4).reshape(d.T, nb, Mr, width)[:, :, :DG]
    return y.reshape(d.T, nb * DG, width)[:, :d.N].contiguous()


def mab_plan_info(d: MabDesc) -> dict:
    """The sharding plan for a 148-SM device (host-only query)."""
    buf = (C.c_int32 * (14 + 6 * 16))()
    _check(lib().lsthm_mab_plan_info(C.byref(d), buf, len(buf)), "lsthm_mab_plan_info")
    keys = ("G", "nr", "DG", "Mr", "ngroups", "nblocks", "cd", "blob_f", "blob_b", "act_f", "act_b", "smem_fwd", "smem_bwd", "ws_group")
    out = dict(zip(keys, list(buf[:14])))
    out["ranks"] = [dict(zip(("m", "u0", "nu", "head", "j0", "nj"), list(buf[14 + 6 * r:20 + 6 * r]))) for r in range(out["G"])]
    return out


# ------------------------------------------------------------------------------------------------
# lsthm_sps speaker-state cell
# ------------------------------------------------------------------------------------------------
def _int_ptr(t: torch.Tensor, name: str) -> int:
    if not t.is_cuda or t.dtype != torch.int32 or not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous CUDA int32 tensor")
    return t.data_ptr()


def make_sps_desc(T: int, N: int, rows_per_cta: int = 0, att_p: float = 0.0, att_seed: int = 0) -> SpsDesc:
    d = SpsDesc()
    d.T, d.N, d.rows_per_cta, d.att_p, d.att_seed = T, N, rows_per_cta, att_p, att_seed
    return d


def make_sps_weights(U, V, S, Wih, Whh, bq, Wq, Wk) -> SpsWeights:
    w = SpsWeights()
    for c in range(2):
        w.U[c], w.V[c], w.S[c] = _dev_ptr(U[c], "U"), _dev_ptr(V[c], "V"), _dev_ptr(S[c], "S")
        w.Wih[c], w.Whh[c], w.bq[c] = _dev_ptr(Wih[c], "Wih"), _dev_ptr(Whh[c], "Whh"), _dev_ptr(bq[c], "bq")
    w.Wq, w.Wk = _dev_ptr(Wq, "Wq"), _dev_ptr(Wk, "Wk")
    return w


def make_sps_masks(mq0=None, mq1=None, ml=None, ma=None, att_mask=None) -> SpsMasks:
    m = SpsMasks()
    m.mq[0], m.mq[1] = _dev_ptr(mq0, "mq0"), _dev_ptr(mq1, "mq1")
    m.ml, m.ma, m.att_mask = _dev_ptr(ml, "ml"), _dev_ptr(ma, "ma"), _dev_ptr(att_mask, "att_mask")
    return m


def sps_packed_floats() -> int:
    return lib().lsthm_sps_packed_floats()


def sps_workspace_floats(d: SpsDesc) -> int:
    n = lib().lsthm_sps_workspace_floats(C.byref(d))
    if n == 0:
        raise RuntimeError(f"bad sps descriptor: {lib().lsthm_last_error().decode()}")
    return n


def sps_pack(w: SpsWeights, packed: torch.Tensor) -> None:
    _check(lib().lsthm_sps_pack(C.byref(w), _dev_ptr(packed, "packed"), _stream()), "lsthm_sps_pack")


def sps_fwd(d, w, packed, gx, qmask, pi, n0, masks, workspace, out, sGQ, sCQ, sHQ, sXQ, sGL, sCL, sHL) -> None:
    _check(lib().lsthm_sps_fwd(C.byref(d), C.byref(w), _dev_ptr(packed, "packed"), _dev_ptr(gx, "gx"),
                               _dev_ptr(qmask, "qmask"), _int_ptr(pi, "pi"), _int_ptr(n0, "n0"), C.byref(masks),
                               _dev_ptr(workspace, "workspace"), _dev_ptr(out, "out"), _dev_ptr(sGQ, "sGQ"),
                               _dev_ptr(sCQ, "sCQ"), _dev_ptr(sHQ, "sHQ"), _dev_ptr(sXQ, "sXQ"), _dev_ptr(sGL, "sGL"),
                               _dev_ptr(sCL, "sCL"), _dev_ptr(sHL, "sHL"), _stream()), "lsthm_sps_fwd")


def sps_bwd(d, w, qmask, pi, pr, n0, masks, dout, sGQ, sCQ, sGL, sCL, workspace, dGL, dGQ, dWqk) -> None:
    _check(lib().lsthm_sps_bwd(C.byref(d), C.byref(w), _dev_ptr(qmask, "qmask"), _int_ptr(pi, "pi"), _int_ptr(pr, "pr"),
                               _int_ptr(n0, "n0"), C.byref(masks), _dev_ptr(dout, "dout"), _dev_ptr(sGQ, "sGQ"),
                               _dev_ptr(sCQ, "sCQ"), _dev_ptr(sGL, "sGL"), _dev_ptr(sCL, "sCL"),
                               _dev_ptr(workspace, "workspace"), _dev_ptr(dGL, "dGL"), _dev_ptr(dGQ, "dGQ"),
                               _dev_ptr(dWqk, "dWqk"), _stream()), "lsthm_sps_bwd")


def sps_launch_info(d: SpsDesc) -> dict:
    v = [C.c_int32() for _ in range(5)]
    _check(lib().lsthm_sps_launch_info(C.byref(d), *[C.byref(x) for x in v]), "lsthm_sps_launch_info")
    return dict(zip(("grid", "block", "rows", "smem_fwd", "smem_bwd"), (x.value for x in v)))


# ------------------------------------------------------------------------------------------------
# lsthm_onlysp / lsthm_nsps GRU speaker-state cell
# ------------------------------------------------------------------------------------------------
def make_gsp_desc(T: int, N: int, listener: int, rows_per_cta: int = 0, att_p: float = 0.0, att_seed: int = 0) -> GspDesc:
    d = GspDesc()
    d.T, d.N, d.rows_per_cta, d.listener, d.att_p, d.att_seed = T, N, rows_per_cta, listener, att_p, att_seed
    return d


def make_gsp_weights(U, V, S, Whh, bhh, Wq, Wk) -> GspWeights:
    w = GspWeights()
    for c in range(2):
        w.U[c], w.V[c], w.S[c] = _dev_ptr(U[c], "U"), _dev_ptr(V[c], "V"), _dev_ptr(S[c], "S")
    w.Whh, w.bhh = _dev_ptr(Whh, "Whh"), _dev_ptr(bhh, "bhh")
    w.Wq, w.Wk = _dev_ptr(Wq, "Wq"), _dev_ptr(Wk, "Wk")
    return w


def make_gsp_masks(ms=None, ml=None, ma=None, att_mask=None) -> GspMasks:
    m = GspMasks()
    m.ms, m.ml, m.ma, m.att_mask = _dev_ptr(ms, "ms"), _dev_ptr(ml, "ml"), _dev_ptr(ma, "ma"), _dev_ptr(att_mask, "att_mask")
    return m


def gsp_packed_floats() -> int:
    return lib().lsthm_gsp_packed_floats()


def gsp_pack(w: GspWeights, packed: torch.Tensor) -> None:
    _check(lib().lsthm_gsp_pack(C.byref(w), _dev_ptr(packed, "packed"), _stream()), "lsthm_gsp_pack")


def gsp_fwd(d, w, packed, gx, gxs, qmask, masks, out, sGS, sQS, sGL, sCL) -> None:
    _check(lib().lsthm_gsp_fwd(C.byref(d), C.byref(w), _dev_ptr(packed, "packed"), _dev_ptr(gx, "gx"), _dev_ptr(gxs, "gxs"),
                               _dev_ptr(qmask, "qmask"), C.byref(masks), _dev_ptr(out, "out"), _dev_ptr(sGS, "sGS"),
                               _dev_ptr(sQS, "sQS"), _dev_ptr(sGL, "sGL"), _dev_ptr(sCL, "sCL"), _stream()), "lsthm_gsp_fwd")


def gsp_bwd(d, w, qmask, masks, dout, sGS, sQS, sGL, sCL, dGL, dGi, dGh, dWqk) -> None:
    _check(lib().lsthm_gsp_bwd(C.byref(d), C.byref(w), _dev_ptr(qmask, "qmask"), C.byref(masks), _dev_ptr(dout, "dout"),
                               _dev_ptr(sGS, "sGS"), _dev_ptr(sQS, "sQS"), _dev_ptr(sGL, "sGL"), _dev_ptr(sCL, "sCL"),
                               _dev_ptr(dGL, "dGL"), _dev_ptr(dGi, "dGi"), _dev_ptr(dGh, "dGh"), _dev_ptr(dWqk, "dWqk"),
                               _stream()), "lsthm_gsp_bwd")


def gsp_launch_info(d: GspDesc) -> dict:
    v = [C.c_int32() for _ in range(5)]
    _check(lib().lsthm_gsp_launch_info(C.byref(d), *[C.byref(x) for x in v]), "lsthm_gsp_launch_info")
    return dict(zip(("grid", "block", "rows", "smem_fwd", "smem_bwd"), (x.value for x in v)))


# ------------------------------------------------------------------------------------------------
# tcgen05 split-bf16 GEMM (time-parallel products)
# ------------------------------------------------------------------------------------------------
GEMM_NT, GEMM_NN, GEMM_TN, GEMM_NT_RELU = 0, 1, 2, 3
GEMM_BF16 = 0x10
GEMM_X6 = 0x20          # six-term split (24-bit operands): products whose sum cancels structurally (sequence cross attention)
# "fp32": every tensor-core product is the fp32-accurate three-term bf16 split (parity mode, the default).
# "bf16": operands rounded to bf16, one UMMA per k-step (BASELINE.json's bf16 mode; tolerance stated in tests/test_bf16_gpu.py).
PRECISION = os.environ.get("LSTHM_PRECISION", "fp32")


def set_precision(mode: str) -> None:
    global PRECISION
    if mode not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    PRECISION = mode
GEMM3W_MIN_ROWS = int(os.environ.get("LSTHM_GEMM3W_MIN_ROWS", "2048"))


def _mat(t: torch.Tensor, name: str):
    """2-D fp32 CUDA tensor with unit inner stride -> (ptr, leading dimension)."""
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1):
        raise RuntimeError(f"{name}: expected a 2-D float32 CUDA tensor with contiguous rows")
    if t.data_ptr() % 16 or t.stride(0) % 4:
        raise RuntimeError(f"{name}: rows must be 16-byte aligned (leading dimension multiple of 4)")
    _on_current_device(t, name)
    return t.data_ptr(), t.stride(0)


def gemm3(mode: int, a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None,
          out: Optional[torch.Tensor] = None, x6: bool = False) -> torch.Tensor:
    """mode NT: a[M,K] @ b[N,K]^T (+bias);  NN: a[M,K] @ b[K,N];  TN: a[K,M]^T @ b[K,N].
    ``out`` (optional): a 2-D fp32 CUDA view [M, N] with unit inner stride (e.g. a column block of a wider matrix).
    ``x6``: the six-term split (LSTHM_GEMM_X6) regardless of the precision mode."""
    if mode in (GEMM_NT, GEMM_NT_RELU):
        M, K = a.shape; N = b.shape[0]; assert b.shape[1] == K
    elif mode == GEMM_NN:
        M, K = a.shape; N = b.shape[1]; assert b.shape[0] == K
    else:
        K, M = a.shape; N = b.shape[1]; assert b.shape[0] == K
    pa, lda = _mat(a, "a")
    pb, ldb = _mat(b, "b")
    if out is None:
        c, ldc = torch.empty(M, N, device=a.device, dtype=torch.float32), N
    else:
        if tuple(out.shape) != (M, N):
            raise RuntimeError(f"gemm3: out has shape {tuple(out.shape)}, expected {(M, N)}")
        c = out
        _, ldc = _mat(out, "out")
    base_mode = mode
    if x6:
        mode |= GEMM_X6
    elif PRECISION == "bf16":
        mode |= GEMM_BF16
    if base_mode != GEMM_TN and M >= GEMM3W_MIN_ROWS and not x6:
        # B is a layer weight and there are many rows: pre-split weight images + 128 x 256 tiles
        nbytes = lib().lsthm_gemm3w_pack_bytes(N, K)
        pack = torch.empty(nbytes, device=a.device, dtype=torch.uint8)
        _check(lib().lsthm_gemm3w(mode, M, N, K, pa, lda, pb, ldb, _dev_ptr(bias, "bias"), c.data_ptr(), ldc, pack.data_ptr(), nbytes,
                                  _stream()), "lsthm_gemm3w")
        return c
    nws = lib().lsthm_gemm3_workspace_floats(mode, M, N, K)
    ws = torch.empty(nws, device=a.device, dtype=torch.float32) if nws else None
    _check(lib().lsthm_gemm3(mode, M, N, K, pa, lda, pb, ldb, _dev_ptr(bias, "bias"), c.data_ptr(), ldc,
                             None if ws is None else ws.data_ptr(), nws, _stream()), "lsthm_gemm3")
    return c


# ------------------------------------------------------------------------------------------------
# fused encoder self-attention (tcgen05)
# ------------------------------------------------------------------------------------------------
def make_attn_desc(B, L, H, ldq, ldk, ldv, ldo, scale, p_drop=0.0, seed=0, d_head=40, time_major=False) -> AttnDesc:
    d = AttnDesc()
    d.B, d.L, d.H, d.d_head, d.ldq, d.ldk, d.ldv, d.ldo = B, L, H, d_head, ldq, ldk, ldv, ldo
    d.scale, d.p_drop, d.seed = scale, p_drop, seed
    d.row_stride_b, d.row_stride_i = (1, B) if time_major else (L, 1)
    d.precision = 1 if PRECISION == "bf16" else 0
    return d


def _f32_cuda(t: torch.Tensor, name: str) -> int:
    if not (t.is_cuda and t.dtype == torch.float32 and t.data_ptr() % 16 == 0):
        raise RuntimeError(f"{name}: expected a 16-byte aligned float32 CUDA tensor")
    _on_current_device(t, name)
    return t.data_ptr()


def attn_fwd(d: AttnDesc, q, k, v, out, lse) -> None:
    _check(lib().lsthm_attn_fwd(C.byref(d), _f32_cuda(q, "q"), _f32_cuda(k, "k"), _f32_cuda(v, "v"), _f32_cuda(out, "out"),
                                _f32_cuda(lse, "lse"), _stream()), "lsthm_attn_fwd")


def attn_bwd(d: AttnDesc, q, k, v, out, lse, dout, dq, dk, dv) -> None:
    _check(lib().lsthm_attn_bwd(C.byref(d), _f32_cuda(q, "q"), _f32_cuda(k, "k"), _f32_cuda(v, "v"), _f32_cuda(out, "out"),
                                _f32_cuda(lse, "lse"), _f32_cuda(dout, "dout"), _f32_cuda(dq, "dq"), _f32_cuda(dk, "dk"), _f32_cuda(dv, "dv"),
                                _stream()), "lsthm_attn_bwd")


def make_xattn_desc(B, L, D, ldq, ldk, ldv, ldo, scale, p_drop=0.0, seed=0, time_major=True) -> XAttnDesc:
    d = XAttnDesc()
    d.B, d.L, d.D, d.ldq, d.ldk, d.ldv, d.ldo = B, L, D, ldq, ldk, ldv, ldo
    d.scale, d.p_drop, d.seed = scale, p_drop, seed
    d.row_stride_b, d.row_stride_i = (1, B) if time_major else (L, 1)
    return d


def xattn_fwd(d: XAttnDesc, q, k, v, out, lse=None) -> None:
    _check(lib().lsthm_xattn_fwd(C.byref(d), _f32_cuda(q, "q"), _f32_cuda(k, "k"), _f32_cuda(v, "v"), _f32_cuda(out, "out"),
                                 None if lse is None else _f32_cuda(lse, "lse"), _stream()), "lsthm_xattn_fwd")


def xattn_bwd(d: XAttnDesc, q, k, v, dout, dq, dk, dv) -> None:
    _check(lib().lsthm_xattn_bwd(C.byref(d), _f32_cuda(q, "q"), _f32_cuda(k, "k"), _f32_cuda(v, "v"), _f32_cuda(dout, "dout"),
                                 _f32_cuda(dq, "dq"), _f32_cuda(dk, "dk"), _f32_cuda(dv, "dv"), _stream()), "lsthm_xattn_bwd")


def reverse_seq(X: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
    """X [L,B,w] fp32 contiguous, lens [B] int32 -> per-dialogue flipped copy (lsthm_sps.py:396-410)."""
    L_, B, w = X.shape
    out = torch.empty_like(X)
    _check(lib().lsthm_reverse_seq(L_, B, w, _dev_ptr(X, "X"), _int_ptr(lens, "lens"), out.data_ptr(), _stream()), "lsthm_reverse_seq")
    return out


def _i64_ptr(t: torch.Tensor, name: str) -> int:
    if not (t.is_cuda and t.dtype == torch.int64 and t.is_contiguous()):
        raise RuntimeError(f"{name}: expected a contiguous CUDA int64 tensor")
    _on_current_device(t, name)
    return t.data_ptr()


def masked_loss_fwd(kind: int, pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """-> out2 = [loss, sum(mask)] (device)."""
    R, Cn = pred.shape
    ws = torch.empty(lib().lsthm_masked_loss_workspace_floats(R), device=pred.device, dtype=torch.float32)
    out2 = torch.empty(2, device=pred.device, dtype=torch.float32)
    _check(lib().lsthm_masked_loss_fwd(R, Cn, kind, _dev_ptr(pred, "pred"), _i64_ptr(target, "target"), _dev_ptr(mask, "mask"),
                                       ws.data_ptr(), out2.data_ptr(), _stream()), "lsthm_masked_loss_fwd")
    return out2


def masked_loss_bwd(kind: int, pred, target, mask, out2, gout) -> torch.Tensor:
    R, Cn = pred.shape
    dpred = torch.empty_like(pred)
    _check(lib().lsthm_masked_loss_bwd(R, Cn, kind, _dev_ptr(pred, "pred"), _i64_ptr(target, "target"), _dev_ptr(mask, "mask"),
                                       out2.data_ptr(), _dev_ptr(gout, "gout"), dpred.data_ptr(), _stream()), "lsthm_masked_loss_bwd")
    return dpred


def _rows2d(t: torch.Tensor, name: str):
    """(pointer, row stride) of a 2-D fp32 CUDA matrix with unit inner stride and 16-byte aligned rows."""
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 4 == 0
            and t.data_ptr() % 16 == 0):
        raise RuntimeError(f"{name}: expected a 2-D float32 CUDA matrix with unit inner stride and 16-byte aligned rows")
    return t.data_ptr(), t.stride(0)


def make_dln_desc(R, d, eps, p_drop=0.0, seed=0) -> DlnDesc:
    x = DlnDesc()
    x.R, x.d, x.eps, x.p_drop, x.seed = R, d, eps, p_drop, seed
    return x


def dln_fwd(desc: DlnDesc, y, bias, res, gamma, beta, v, out) -> None:
    (py, ldy), (pr, ldr), (po, ldo) = _rows2d(y, "y"), _rows2d(res, "res"), _rows2d(out, "out")
    pv, ldv = _rows2d(v, "v") if v is not None else (None, 0)
    _check(lib().lsthm_dln_fwd(C.byref(desc), py, ldy, _dev_ptr(bias, "bias"), pr, ldr, _dev_ptr(gamma, "gamma"), _dev_ptr(beta, "beta"), pv, ldv, po, ldo,
                               _stream()), "lsthm_dln_fwd")


def dln_bwd(desc: DlnDesc, dout, v, gamma, dy, dres, dgamma, dbeta, dbias) -> None:
    (pdo, lddo), (pv, ldv), (pdr, lddr) = _rows2d(dout, "dout"), _rows2d(v, "v"), _rows2d(dres, "dres")
    pdy, lddy = _rows2d(dy, "dy") if dy is not None else (None, 0)
    nws = lib().lsthm_dln_workspace_floats(desc.d)
    ws = torch.empty(nws, device=dout.device, dtype=torch.float32)
    _check(lib().lsthm_dln_bwd(C.byref(desc), pdo, lddo, pv, ldv, _dev_ptr(gamma, "gamma"), pdy, lddy, pdr, lddr,
                               _dev_ptr(dgamma, "dgamma"), _dev_ptr(dbeta, "dbeta"), _dev_ptr(dbias, "dbias"), ws.data_ptr(), nws,
                               _stream()), "lsthm_dln_bwd")


def colsum(a: torch.Tensor) -> torch.Tensor:
    """Column sums of a 2-D fp32 CUDA matrix (unit inner stride, 16-byte aligned rows, width % 4 == 0)."""
    pa, ld = _rows2d(a, "a")
    R, Cc = a.shape
    out = torch.empty(Cc, device=a.device, dtype=torch.float32)
    nws = lib().lsthm_colsum_workspace_floats(R, Cc)
    ws = torch.empty(nws, device=a.device, dtype=torch.float32)
    _check(lib().lsthm_colsum(R, Cc, pa, ld, out.data_ptr(), ws.data_ptr(), nws, _stream()), "lsthm_colsum")
    return out


def assemble_input(r1, r2, r3, r4, acouf) -> torch.Tensor:
    """x = cat((r1 + r2 + r3 + r4) / 4, acouf) over the last dim (model_trainer.py:104-105), one fused pass."""
    lead, dt, da = r1.shape[:-1], r1.shape[-1], acouf.shape[-1]
    for t in (r2, r3, r4):
        if t.shape != r1.shape:
            raise RuntimeError("assemble_input: the four text layers must have the same shape")
    if acouf.shape[:-1] != lead:
        raise RuntimeError("assemble_input: acouf must share the leading dims of the text layers")
    x = torch.empty(*lead, dt + da, device=r1.device, dtype=torch.float32)
    R = x.numel() // (dt + da)
    _check(lib().lsthm_assemble_input(R, dt, da, _dev_ptr(r1, "r1"), _dev_ptr(r2, "r2"), _dev_ptr(r3, "r3"), _dev_ptr(r4, "r4"),
                                      _dev_ptr(acouf, "acouf"), x.data_ptr(), _stream()), "lsthm_assemble_input")
    return x


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step) -> None:
    """Fused Adam on flat, contiguous fp32 CUDA buffers of equal length (in place)."""
    n = param.numel()
    for t in (grad, exp_avg, exp_avg_sq):
        if t.numel() != n:
            raise RuntimeError("adam_step: buffers must have the same length")
    _check(lib().lsthm_adam_step(_dev_ptr(param, "param"), _dev_ptr(grad, "grad"), _dev_ptr(exp_avg, "exp_avg"),
                                 _dev_ptr(exp_avg_sq, "exp_avg_sq"), n, lr, beta1, beta2, eps, weight_decay, step, _stream()),
           "lsthm_adam_step")
