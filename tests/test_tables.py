"""Derived Ozaki-II constants (gemmul8_b200/tables.py) == the reference's literal tables
(GEMMul8/src/table.hpp, parsed into tests/golden/ref_tables.json by tools/extract_ref_tables.py)."""
import json
import subprocess
import sys
from math import gcd
from pathlib import Path

import pytest

from gemmul8_b200 import tables as T

ROOT = Path(__file__).resolve().parent.parent
G = json.loads((ROOT / "tests/golden/ref_tables.json").read_text())


@pytest.mark.parametrize("be", ["INT8", "FP8"])
def test_moduli_pairwise_coprime_and_match(be):
    mods = T.moduli(be)
    assert list(mods) == G[be]["moduli"]
    for i in range(len(mods)):
        for j in range(i):
            assert gcd(mods[i], mods[j]) == 1


@pytest.mark.parametrize("be", ["INT8", "FP8"])
def test_P_invP_log2P(be):
    for n in range(2, 21):
        hi, lo = T.P_dd(be, n)
        assert [hi.hex(), lo.hex()] == G[be]["P"][n - 2]
        assert T.invP(be, n).hex() == G[be]["invP"][n - 2]
        assert T.log2P(be, n).hex() == G[be]["log2P"][str(n)]


@pytest.mark.parametrize("be", ["INT8", "FP8"])
def test_crt_weights(be):
    pd = T.THRESHOLD[be]["P_is_double"]
    for n in range(2, 21):
        ws = T.crt_weights(be, n)
        for w, p in zip(ws, T.moduli(be)):
            assert w % p == 1 and all(w % q == 0 for q in T.moduli(be)[:n] if q != p)
        assert [x.hex() for x in T.qPi_1(be, n)] == G[be]["qPi_1"][n - 2][:n]
        if n > pd:
            assert [[a.hex(), b.hex()] for a, b in T.qPi_2(be, n)] == G[be]["qPi_2"][n - pd - 1][:n]
            # the hi chain must be exact: all hi parts share one quantum and the worst-case sum stays below 2^53 quanta
            his = [int(a) for a, _ in T.qPi_2(be, n)]
            q = min((h & -h) for h in his if h)
            assert sum((h // q) * (p // 2) for h, p in zip(his, T.moduli(be))) < 2 ** 53


@pytest.mark.parametrize("be", ["INT8", "FP8"])
def test_mod_pow2(be):
    assert T.mod_pow2(be) == G[be]["mod_pow2"]


def test_num_mat():
    assert [T.num_mat("INT8", n) for n in (2, 6, 14, 20)] == [2, 6, 14, 20]
    assert [T.num_mat("FP8", n) for n in (2, 6, 7, 14, 20)] == [4, 12, 15, 36, 54]
    assert G["sqrt_moduli"][1:] == list(T.FP8_SQRT_MODULI) or G["sqrt_moduli"][-6:] == list(T.FP8_SQRT_MODULI)


def test_generated_header_is_current():
    pytest.importorskip("mpmath")
    r = subprocess.run([sys.executable, str(ROOT / "tools/gen_tables.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
