"""The C-ABI library loads on a machine WITHOUT a GPU and exports every symbol include/gemmul8_c.h declares;
host-only entry points behave like the reference (workSize golden values from the reference library)."""
import ctypes
import json
import re
from pathlib import Path

import pytest

from gemmul8_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    hdr = (ROOT / "include/gemmul8_c.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(g8_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_all_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gemmul8_c.h but not exported by libg8core.so"
    assert set(names) == set(_lib.SYMBOLS), "python binding table out of sync with the header"


def test_version_string():
    assert b"sm_100a" in _lib.load().g8_version()


def test_work_size_matches_reference_golden():
    import gemmul8_b200 as g8

    gold = json.loads((ROOT / "tests/golden/worksize.json").read_text())["rows"]
    for cplx, be, m, n, k, N, ea, eb, tot, wa, wb in gold:
        assert g8.work_size(m, n, k, N, bool(cplx), be, bool(ea), bool(eb)) == (tot, wa, wb), (cplx, be, m, n, k, N, ea, eb)


def test_compute_entry_points_fail_loudly_without_gpu():
    """No CPU fallback: without a CUDA device the compute calls return an error code, never a result."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    d = _lib.GemmDesc()
    d.dtype, d.backend, d.m, d.n, d.k, d.num_moduli = 1, 0, 4, 4, 4, 14
    buf = (ctypes.c_double * 64)()
    p = ctypes.addressof(buf)
    d.A = d.B = d.C = d.alpha = d.beta = d.work = p
    assert lib.g8_gemm(ctypes.byref(d), None) != 0
    d.num_moduli = 1
    assert lib.g8_gemm(ctypes.byref(d), None) == 10001  # INVALID_VALUE before touching the device


def test_native_library_missing_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setenv("GEMMUL8_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.NativeLibraryMissing):
        _lib.load()


def test_cxx_and_hook_symbols_match_reference():
    """lib/libgemmul8.so exports the reference library's 16 C++ symbols (nm of the compiled reference,
    tests/golden/ref_cxx_symbols.txt) and the 6 hook symbols (hook.cu:846-1055)."""
    import subprocess

    lib = ROOT / "gemmul8_b200" / "lib" / "libgemmul8.so"
    assert lib.exists(), "run __graft_entry__.build() first"
    want = (ROOT / "tests/golden/ref_cxx_symbols.txt").read_text().split()
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib)], capture_output=True, text=True).stdout
    have = set(re.findall(r" T (\S+)", out))
    assert len(want) == 16 and set(want) <= have
    assert {"cublasSgemm_v2", "cublasDgemm_v2", "cublasCgemm_v2", "cublasZgemm_v2", "cublasGemmEx", "cublasDestroy_v2"} <= have


def test_argument_validation_needs_no_gpu():
    """Argument errors are reported before the device is touched: the k limits of both backends (2^17 INT8; 2^16 FP8, where binary32
    accumulation of the piece products stops being exact -- ADVICE r01), and the plan / communicator constructors of the host-buffer
    and multi-GPU entry points."""
    lib = _lib.load()
    INVALID = 10001
    d = _lib.GemmDesc()
    buf = (ctypes.c_double * 64)()
    p = ctypes.addressof(buf)
    d.A = d.B = d.C = d.alpha = d.beta = d.work = p
    d.dtype, d.m, d.n, d.num_moduli = 1, 4, 4, 14
    d.backend, d.k = 0, (1 << 17) + 1
    assert lib.g8_gemm(ctypes.byref(d), None) == INVALID
    d.backend, d.k = 1, (1 << 16) + 1
    assert lib.g8_gemm(ctypes.byref(d), None) == INVALID
    plan = ctypes.c_void_p()
    assert lib.g8_host_plan_create(ctypes.byref(plan), 7, 0, 0, 0, 8, 8, 8, 14, 0, 0) == INVALID          # dtype
    assert lib.g8_host_plan_create(ctypes.byref(plan), 1, 0, 0, 3, 8, 8, 8, 14, 0, 0) == INVALID          # op_B
    assert lib.g8_host_plan_create(ctypes.byref(plan), 1, 1, 0, 0, 8, 8, (1 << 16) + 1, 8, 0, 0) == INVALID   # FP8 k limit
    assert lib.g8_host_plan_create(ctypes.byref(plan), 1, 0, 0, 0, 8, 8, 8, 21, 0, 0) == INVALID          # num_moduli
    assert lib.g8_gemm_host(None, p, p, 8, p, 8, p, p, 8, None) == INVALID
    comm = ctypes.c_void_p()
    handle = ctypes.create_string_buffer(64)
    assert lib.g8_mg_comm_create(ctypes.byref(comm), 0, 0, 4096, handle) == INVALID
    assert lib.g8_mg_comm_create(ctypes.byref(comm), 9, 0, 4096, handle) == INVALID                       # more than the 8 GPUs of a box
    assert lib.g8_mg_comm_create(ctypes.byref(comm), 2, 2, 4096, handle) == INVALID                       # rank out of range
    assert lib.g8_mg_plan_create(ctypes.byref(plan), None, 1, 0, 0, 256, 256, 256, 14, 0) == INVALID
    assert lib.g8_mg_plan_create_backend(ctypes.byref(plan), None, 1, 1, 0, 0, 256, 256, 256, 14, 0) == INVALID   # no communicator
    assert lib.g8_mg_plan_create_backend(ctypes.byref(plan), None, 1, 2, 0, 0, 256, 256, 256, 14, 0) == INVALID   # backend
    assert lib.g8_gemm_mg(None, p, p, 8, p, 8, p, p, 8, None) == INVALID
    assert lib.g8_stage_gemm_bound_chain(p, 256, p, 256, 8, 8, 256, 0, p, p, None) != 0                   # chain < 1 (or no device)
