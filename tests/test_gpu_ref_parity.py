"""Bitwise parity with the UNMODIFIED reference library (oracle/_ref/libgemmul8_ref.so, built from
/root/reference by oracle/Makefile) on the same B200: shift exponents, C_mid residues and C must be identical for
S/D/C/ZGEMM, fast and accurate mode.  This is what pins the parts a CPU cannot restate (MUFU.LG2, RU sum order)."""
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_bitwise_parity_with_reference_library(cuda):
    so = ROOT / "oracle/_ref/libgemmul8_ref.so"
    if not so.exists():
        pytest.skip("oracle/_ref/libgemmul8_ref.so not built (needs /root/reference at build time)")
    sys.path.insert(0, str(ROOT / "tools"))
    import gpu_debug as D

    D.FAILS.clear()
    D.check_ref()
    assert not D.FAILS, D.FAILS


def test_odd_shapes_bit_identical_to_reference(cuda):
    """tools/odd_shapes.py: n beyond the reference's 12288-column chunking, tall-skinny, k = 2^17 (INT8) / 65536 (FP8), sizes that are
    multiples of no tile dimension, all four types, both backends: the complete C must equal the reference library's bit for bit."""
    import subprocess

    so = ROOT / "oracle/_ref/libgemmul8_ref.so"
    if not so.exists():
        pytest.skip("oracle/_ref/libgemmul8_ref.so not built (needs /root/reference at build time)")
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "odd_shapes.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ODD SHAPES: all bit-identical" in r.stdout, r.stdout[-3000:]
