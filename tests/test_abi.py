"""CPU tests of the C-ABI boundary: the shared library builds (nvcc cross-compiles without a GPU),
loads, and exports exactly the entry points include/baybe_b200.h declares; the ctypes mirrors of
the C structs have the C compiler's sizes.  No compute calls (no GPU here)."""
from __future__ import annotations

import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "baybe_b200.h"


@pytest.fixture(scope="module")
def lib():
    from baybe_b200 import _lib
    from baybe_b200.build import build

    build(verbose=False)
    return _lib.load()


def _declared_functions() -> set[str]:
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return set(re.findall(r"\b(bb_[a-z0-9_]+)\s*\(", text))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from baybe_b200 import _lib

    declared = _declared_functions()
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True,
                         check=True).stdout
    exported = set(re.findall(r" T (bb_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    assert lib.bb_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_the_c_compiler(tmp_path):
    from baybe_b200 import _lib

    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include "baybe_b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n",'
        "sizeof(bb_model_desc),sizeof(bb_model),sizeof(bb_acq_spec),sizeof(bb_best));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(HEADER.parent), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.ModelDesc), ctypes.sizeof(_lib.Model), ctypes.sizeof(_lib.AcqSpec),
                     ctypes.sizeof(_lib.Best)]


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "c89ish.c"
    src.write_text('#include "baybe_b200.h"\nint main(void){return sizeof(bb_model) == 0;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", str(HEADER.parent), str(src)],
                   check=True)


def test_status_codes_and_blob_size_queries_without_gpu(lib):
    assert lib.bb_model_blob_bytes(0, 3, 1) == 0
    b1 = lib.bb_model_blob_bytes(256, 20, 1)
    b2 = lib.bb_model_blob_bytes(512, 20, 1)
    assert 0 < b1 < b2
    # argument validation happens before any CUDA call
    from baybe_b200 import _lib

    rc = lib.bb_acq_score(None, None, None, 10, None, 0, None, None)
    assert rc == _lib.BB_ERR_INVALID and b"null" in lib.bb_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "bb_acq_score")


def test_round2_entry_points_validate_their_arguments_before_any_cuda_call(lib):
    import ctypes as C

    from baybe_b200 import _lib

    # bb_nei_reduce: the caller's GEMM buffer must hold S sample columns + m root columns per row
    rc = lib.bb_nei_reduce(None, 10, 8, 4, None, None, None, None, C.c_float(1.0), C.c_float(0.0), 5, None, None)
    assert rc == _lib.BB_ERR_INVALID and b"bad shape" in lib.bb_last_error()
    assert lib.bb_nei_reduce(None, 12, 8, 4, None, None, None, None, C.c_float(1.0), C.c_float(0.0), 0, None, None) == 0
    rc = lib.bb_nei_reduce(None, 12, 8, 4, None, None, None, None, C.c_float(1.0), C.c_float(0.0), 5, None, None)
    assert rc == _lib.BB_ERR_INVALID and b"null" in lib.bb_last_error()
    # bb_score_fused_overlapped: model / acquisition spec are checked first
    rc = lib.bb_score_fused_overlapped(None, None, None, 0, 10, 20, None, 0, None, 0, None, None, None, None, 0, None,
                                       None, 0, None, None)
    assert rc == _lib.BB_ERR_INVALID and b"missing" in lib.bb_last_error()
    # bb_allreduce_best: group sanity
    g = _lib.PeerGroup()
    g.rank, g.world = 3, 2
    rc = lib.bb_allreduce_best(C.byref(g), C.c_void_p(8), 0, C.c_void_p(8), C.c_void_p(8), None)
    assert rc == _lib.BB_ERR_INVALID and b"rank 3" in lib.bb_last_error()


def test_product_has_no_cpu_fallback():
    import torch

    from baybe_b200 import DeviceGP
    from baybe_b200.synthetic import mixed_small_workload

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeviceGP(**mixed_small_workload().gp_kwargs())


def test_product_never_imports_the_oracle():
    for py in (ROOT / "baybe_b200").rglob("*.py"):
        text = py.read_text()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), py
