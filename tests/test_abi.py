"""CPU: the C-ABI library builds, loads and exports every symbol include/hf_b200.h declares."""
import ctypes
import os
import re

from helpers import ROOT

from pytorchhessianfree_b200 import _lib
from pytorchhessianfree_b200 import build as hf_build


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hf_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = hf_build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in hf_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table and header disagree"


def test_no_compute_entry_points_need_a_gpu_to_load():
    lib = _lib.load()
    assert lib.hf_abi_version() == 2
    assert lib.hf_pcg_state_bytes(250) > 250 * 8
    assert lib.hf_pcg_m_iters_offset() % 8 == 0
    assert ctypes.sizeof(_lib.PcgStatus) == 80 and ctypes.sizeof(_lib.LayerDesc) == 88


def test_product_code_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "pytorchhessianfree_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "hf_oracle" not in text and "oracle/" not in text, f"{f} references the oracle"
