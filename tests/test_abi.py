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
    """The group plan is a pure host computation (148-SM device assumed when there is no GPU)."""
    d = lib.make_desc(110, 1024, (128, 16, 64), (16, 128, 100))
    assert lib.mab_pack_bytes(d) > 2 * 1450000 and lib.mab_workspace_bytes(d) > 0
    info = lib.mab_launch_info(d)
    plan = lib.mab_plan_info(d)
    # ATV: 13 ranks per group (8 + 1 + 4 slices of 16 hidden units; 4 heads x 3 feature ranges), 11 co-resident groups
    assert info["group"] == 13 and info["grid"] == 143 and info["block"] == 384 and info["dialogues_per_group"] == 96
    assert info["smem_fwd"] <= 227 * 1024 and info["smem_bwd"] <= 227 * 1024 and info["padded_rows"] >= 1024
    ranks = plan["ranks"]
    assert sorted((r["u0"], r["nu"]) for r in ranks) == [(16 * i, 16) for i in range(13)]          # every hidden unit owned once
    slices = sorted((r["head"], r["j0"], r["nj"]) for r in ranks if r["head"] >= 0)
    assert slices == [(k, j0, nj) for k in range(4) for j0, nj in ((0, 64), (64, 64), (128, 80))]   # every logit owned once
    assert [r["m"] for r in ranks] == sorted(r["m"] for r in ranks)                                 # modality ranks contiguous
    at = lib.mab_launch_info(lib.make_desc(110, 32, (128, 16), (16, 128)))
    assert at["group"] == 9 and at["dialogues_per_group"] == 8 and at["grid"] == 36
    for rows in (8, 16, 48, 96, 200):
        i8 = lib.mab_launch_info(lib.make_desc(110, 1024, (128, 16, 64), (16, 128, 100), rows_per_cta=rows))
        assert i8["dialogues_per_group"] == min(rows, 96) and max(i8["smem_fwd"], i8["smem_bwd"]) <= 227 * 1024
    with pytest.raises(RuntimeError, match="multiples of 16"):
        lib.mab_pack_bytes(lib.make_desc(4, 4, (130, 16), (16, 128)))
    with pytest.raises(RuntimeError, match="n_att"):
        lib.mab_pack_bytes(lib.make_desc(4, 4, (128, 16), (16, 128), n_att=2))
    with pytest.raises(RuntimeError, match="map_h"):
        lib.mab_pack_bytes(lib.make_desc(4, 4, (128, 16), (16, 128), map_h=32))


def test_binding_rejects_host_tensors():
    d = lib.make_desc(2, 2, (128, 16), (16, 128))
    with pytest.raises(RuntimeError, match="CUDA"):
        lib.mab_fwd(d, torch.zeros(256, dtype=torch.uint8), torch.zeros(2, 2, 576), None, torch.zeros(2, 2, 288),
                    torch.zeros(2, 2, 64), None, None, None, None, None, None, torch.zeros(256, dtype=torch.uint8))
