"""CPU suite, part 3: the C-ABI shared library loads and exports every symbol that
include/lsthm_b200.h declares; host-only entry points behave (no compute without a GPU)."""
import ctypes as C
import os
import re
from importlib import import_module

import pytest
import torch

import lsthm_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = import_module(lsthm_b200.__name__ + "._lib")


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lsthm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lsthm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(lib.SO_PATH), "liblsthm_b200.so not built: run __graft_entry__.build()"
    dll = C.CDLL(lib.SO_PATH)
    syms = declared_symbols()
    assert {"lsthm_mab_fwd", "lsthm_mab_bwd", "lsthm_mab_pack", "lsthm_abi_version"} <= set(syms)
    for s in syms:
        assert hasattr(dll, s), f"{s} declared in include/lsthm_b200.h but not exported"
    assert dll.lsthm_abi_version() == lib.ABI_VERSION


def test_layout_queries_and_errors():
    d = lib.make_desc(110, 1024, (128, 16, 64), (16, 128, 100))
    assert lib.mab_packed_floats(d) > 400000
    info = lib.mab_launch_info(d)
    assert info["rows"] == 7 and info["grid"] == 147 and info["block"] == 416
    assert info["smem_fwd"] <= 227 * 1024 and info["smem_bwd"] <= 227 * 1024
    at = lib.mab_launch_info(lib.make_desc(110, 32, (128, 16), (16, 128)))
    assert at["rows"] == 1 and at["grid"] == 32 and at["block"] == 288
    for rows in range(1, 9):
        i8 = lib.mab_launch_info(lib.make_desc(110, 1024, (128, 16, 64), (16, 128, 100), rows_per_cta=rows))
        assert i8["rows"] == rows and max(i8["smem_fwd"], i8["smem_bwd"]) <= 227 * 1024
    with pytest.raises(RuntimeError, match="multiples of 4"):
        lib.mab_packed_floats(lib.make_desc(4, 4, (130, 16), (16, 128)))
    with pytest.raises(RuntimeError, match="n_att"):
        lib.mab_packed_floats(lib.make_desc(4, 4, (128, 16), (16, 128), n_att=2))


def test_binding_rejects_host_tensors():
    d = lib.make_desc(2, 2, (128, 16), (16, 128))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        lib.mab_fwd(d, torch.zeros(8), torch.zeros(2, 2, 576), None, torch.zeros(2, 2, 288), torch.zeros(2, 2, 64),
                    None, None, None)
