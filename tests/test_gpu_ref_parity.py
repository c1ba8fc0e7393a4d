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
