"""Drop-in checks at the reference's own boundary (C++ API + LD_PRELOAD hook), restating its sample/ and
debug/test_hijack.cu programs with real assertions.  The programs in tests/native are compiled against THIS repo's
include/gemmul8.hpp and lib/libgemmul8.so."""
import os
import re
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
NATIVE = ROOT / "tests" / "native"
LIB = ROOT / "gemmul8_b200" / "lib" / "libgemmul8.so"
REF_HOOK = ROOT / "oracle" / "_ref" / "libgemmul8_refhook.so"   # the reference's own lib/libgemmul8.so (gemmul8.cu + hook.cu, unmodified)


@pytest.fixture(scope="module")
def binaries(cuda):
    if not all((NATIVE / "_build" / b).exists() for b in ("sample_kat", "hook_check", "ext_check")):
        subprocess.check_call(["make", "-C", str(NATIVE)])
    return NATIVE / "_build"


def test_cxx_api_sample_known_answer(binaries):
    """sample/dgemm_cuBLAS{,Lt}_int8.cu through gemmul8::gemmLt and gemmul8::gemm: exact product at N=15"""
    r = subprocess.run([str(binaries / "sample_kat")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "exact" in r.stdout, r.stdout + r.stderr


def test_cxx_extension_header_host_and_multi_gpu_front_ends(binaries):
    """include/gemmul8_ext.hpp (HostGemm, MgComm / MgGemm with a world of one rank) linked against lib/libgemmul8.a: bit-identical to
    gemmul8::gemmLt for S/D/C/Z, both backends (tests/native/ext_check.cu)"""
    r = subprocess.run([str(binaries / "ext_check")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ext_check OK" in r.stdout and "MISMATCH" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def _run_hook(binaries, env_extra, preload):
    env = dict(os.environ)
    env.update(env_extra)
    if preload:
        env["LD_PRELOAD"] = str(LIB if preload is True else preload)
    r = subprocess.run([str(binaries / "hook_check")], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    return [(ln.split(" checksum ")[0], float(ln.split(" checksum ")[1])) for ln in r.stdout.splitlines() if " checksum " in ln], r.stderr


def test_ld_preload_hook_matches_native_cublas(binaries):
    native, _ = _run_hook(binaries, {}, False)
    assert len(native) == 13
    # not enabled (NUM_MOD unset -> 0): every call must fall through to the real cuBLAS => identical bits
    passthrough, _ = _run_hook(binaries, {}, True)
    assert passthrough == native
    base = {"GEMMUL8_NUM_MOD_D": "15", "GEMMUL8_NUM_MOD_S": "8", "GEMMUL8_NUM_MOD_Z": "15", "GEMMUL8_NUM_MOD_C": "8"}
    for extra in ({}, {"GEMMUL8_FASTMODE_D": "1", "GEMMUL8_FASTMODE_S": "1", "GEMMUL8_FASTMODE_Z": "1", "GEMMUL8_FASTMODE_C": "1"},
                  {"GEMMUL8_SKIP_SCALE_A": "1", "GEMMUL8_SKIP_SCALE_B": "1", "GEMMUL8_MAX_M": "128", "GEMMUL8_MAX_N": "128",
                   "GEMMUL8_MAX_K": "128", "GEMMUL8_MAX_NUM_MOD": "15"}):
        emu, err = _run_hook(binaries, {**base, **extra}, True)
        assert "failed" not in err, err
        assert [a for a, _ in emu] == [a for a, _ in native]
        for (name, x), (_, y) in zip(emu, native):
            single = name.startswith("sgemm") or name.startswith("cgemm")
            tol = 2e-5 if single else (1e-9 if "FASTMODE_D" in extra else 1e-11)
            assert abs(x - y) <= tol * max(1.0, abs(y)), (name, x, y, extra)
    # GEMMUL8_BACKEND=FP8 (hook.cu:567-584): e4m3 emulation, ~9.2 bits per modulus -> N = 13 / 6 reach the same accuracy class
    emu, err = _run_hook(binaries, {"GEMMUL8_NUM_MOD_D": "13", "GEMMUL8_NUM_MOD_S": "6", "GEMMUL8_NUM_MOD_Z": "13", "GEMMUL8_NUM_MOD_C": "6",
                                    "GEMMUL8_BACKEND": "FP8"}, True)
    assert "failed" not in err, err
    for (name, x), (_, y) in zip(emu, native):
        single = name.startswith("sgemm") or name.startswith("cgemm")
        assert abs(x - y) <= (2e-5 if single else 1e-10) * max(1.0, abs(y)), (name, x, y, "FP8")
    # out-of-range moduli count -> native path again (hook.cu:625-629)
    off, _ = _run_hook(binaries, {"GEMMUL8_NUM_MOD_D": "21", "GEMMUL8_NUM_MOD_S": "14"}, True)
    assert off == native


def test_ld_preload_hook_bit_identical_to_reference_hook(binaries):
    """The same unmodified cuBLAS program under LD_PRELOAD of THIS repo's libgemmul8.so and of the reference's own hook library
    (oracle/_ref/libgemmul8_refhook.so = gemmul8.cu + hook.cu compiled where they lie): every checksum -- S/D/C/ZGEMM, real and
    complex GemmEx, the skip-scaling cache incl. its invalidation, a stream switch, two threads on one handle -- must be EQUAL."""
    if not REF_HOOK.exists():
        pytest.skip("oracle/_ref/libgemmul8_refhook.so not built (needs /root/reference at build time)")
    base = {"GEMMUL8_NUM_MOD_D": "15", "GEMMUL8_NUM_MOD_S": "8", "GEMMUL8_NUM_MOD_Z": "15", "GEMMUL8_NUM_MOD_C": "8"}
    fastm = {"GEMMUL8_FASTMODE_D": "1", "GEMMUL8_FASTMODE_S": "1", "GEMMUL8_FASTMODE_Z": "1", "GEMMUL8_FASTMODE_C": "1"}
    skip = {"GEMMUL8_SKIP_SCALE_A": "1", "GEMMUL8_SKIP_SCALE_B": "1", "GEMMUL8_MAX_M": "128", "GEMMUL8_MAX_N": "128", "GEMMUL8_MAX_K": "128",
            "GEMMUL8_MAX_NUM_MOD": "15"}
    for extra in ({}, fastm, skip, {**fastm, **skip}):
        ours, err1 = _run_hook(binaries, {**base, **extra}, True)
        ref, err2 = _run_hook(binaries, {**base, **extra}, REF_HOOK)
        assert "failed" not in err1, err1
        assert len(ours) == len(ref) == 13
        assert ours == ref, (extra, [(a, x, y) for (a, x), (_, y) in zip(ours, ref) if x != y])


def test_exported_symbols_match_reference():
    want = (ROOT / "tests/golden/ref_cxx_symbols.txt").read_text().split()
    out = subprocess.run(["nm", "-D", "--defined-only", str(LIB)], capture_output=True, text=True).stdout
    have = set(re.findall(r" T (\S+)", out))
    assert set(want) <= have
    assert {"cublasSgemm_v2", "cublasDgemm_v2", "cublasCgemm_v2", "cublasZgemm_v2", "cublasGemmEx", "cublasDestroy_v2"} <= have
