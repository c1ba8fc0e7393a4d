"""Derived Ozaki-II constants (gemmul8_b200/tables.py) == the reference's literal tables
(GEMMul8/src/table.hpp, parsed into tests/golden/ref_tables.json by tools/extract_ref_tables.py)."""
import json
import subprocess
import sys
from math import gcd
from pathlib import Path

import numpy as np
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


@pytest.mark.parametrize("be", ["INT8", "FP8"])
def test_level2_residue_formula_exhaustive(be):
    """Our split's level-2 step (DESIGN 3.1, g8_split.cu:level1/level2): a level-1 remainder r with |r| <= 0.52 M_g becomes
    a = r + B_g (INT8: lifted by + M_g when negative -- replayed as the device's unsigned min), then for every member p of the group
    low = (a * ceil(2^32/p)) mod 2^32, s = hi32(low * p) - h must be THE symmetric residue of r mod p (direct remainder by
    multiply-high).  Checked by brute force over the whole admissible range of every modulus (no sampling)."""
    F = T.fast_mod_tables(be)
    mods = T.moduli(be)
    for gi, members in enumerate(F["groups"]):
        M, B = F["grpM"][gi], F["grp_bias"][gi]
        vmax = int(0.52 * M) + 2
        r = np.arange(-vmax, vmax + 1, dtype=np.int64)
        a = (r + B) & 0xFFFFFFFF                                   # the int32 register, viewed as unsigned
        if F["lift"]:
            a = np.minimum(a, (a + M) & 0xFFFFFFFF)               # VIADDMNMX.U32: min(a, a + M) picks the non-wrapped value
        assert a.max() < 2 ** 31 and np.all((a - (r + B)) % M == 0)
        for i in members:
            p, h, mg = mods[i], F["half"][i], F["magic"][i]
            low = (a * mg) & 0xFFFFFFFF                            # IMAD (low 32 bits)
            s = (((low * p) >> 32) + ((-h) & 0xFFFFFFFF)) & 0xFFFFFFFF   # IMAD.HI.U32 with the addend -h
            s = np.where(s >= 2 ** 31, s - 2 ** 32, s)             # reinterpret as int32
            want = ((r + h) % p) - h                               # symmetric representative in [-h, p - 1 - h]
            assert np.array_equal(s, want), (be, p)
            assert s.min() >= -h and s.max() <= p - 1 - h


@pytest.mark.parametrize("be", ["INT8", "FP8"])
def test_power_of_two_residue_from_group0(be):
    """x mod 256 (INT8) / x mod 1024 (FP8) without a separate reduction: x = r + M_0 q exactly and the low word of
    t = fma(x, 1/M_0, 1.5 2^52) is q mod 2^32, so the low bits of (a_0 - B_0 + M_0 lo32(t)) are those of x (g8_split.cu:level1)."""
    F = T.fast_mod_tables(be)
    M, B = F["grpM"][0], F["grp_bias"][0]
    bits = 8 if be == "INT8" else 10
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.integers(-2 ** 62, 2 ** 62, size=200000), rng.integers(-2 ** 40, 2 ** 40, size=100000), np.arange(-5000, 5000),
                         [2 ** 63 - 2 ** 10, -(2 ** 63 - 2 ** 10)]])
    x = xs.astype(np.float64)                                      # the device holds trunc(a 2^s) as a double
    t = x * (1.0 / M) + T.MAGIC_RINT                               # one rounding like the FMA: |x / M| < 2^51 keeps the product exact enough
    q = t - T.MAGIC_RINT
    # exact replay with Python integers (the numpy line above only picks q; r is computed exactly like the device's FMA does)
    for xv, tv, qv in zip(x[::97], t[::97], q[::97]):
        xi, qi = int(xv), int(qv)
        r = xi - M * qi
        assert abs(r) <= 0.52 * M + 2
        qlo = int(np.float64(tv).view(np.uint64)) & 0xFFFFFFFF     # __double2loint(t)
        a = (r + B) & 0xFFFFFFFF
        z = (a + M * qlo - B) & 0xFFFFFFFF
        assert z % (1 << bits) == xi % (1 << bits), (xi, qi)


@pytest.mark.parametrize("be", ["INT8", "FP8"])
def test_level1_group_remainder_bound(be):
    """level 1: r = x - M rint(x / M) evaluated in binary64 with the 1.5 * 2^52 trick stays an exact integer with |r| <= 0.52 M
    for |x| < 2^63 -- sampled at the extremes (largest x, values next to the rounding boundaries of x / M)."""
    F = T.fast_mod_tables(be)
    rng = np.random.default_rng(0)
    for M in F["grpM"]:
        xs = np.concatenate([rng.integers(-2 ** 62, 2 ** 62, size=20000) * 2 + 1, [2 ** 63 - 2 ** 10, -(2 ** 63 - 2 ** 10)],
                             (np.arange(1, 2000) * M + M // 2).astype(np.int64), (np.arange(1, 2000) * M + M // 2 + 1).astype(np.int64)])
        x = xs.astype(np.float64)                              # the split holds trunc(a 2^s) as a double (<= 53 significant bits)
        xi = x.astype(object)                                  # exact integer values of those doubles
        q = (x * (1.0 / M) + T.MAGIC_RINT) - T.MAGIC_RINT
        r = np.array([float(v) for v in (np.array([int(a) for a in x], dtype=object) - np.array([int(b) for b in q], dtype=object) * M)])
        assert np.all(np.abs(r) <= 0.52 * M + 2), (be, M, np.abs(r).max() / M)


def test_fp8_piece_decomposition_exhaustive():
    """FP8 backend split (g8_split.cu:split_store_f8): every symmetric residue r of every FP8 modulus is written as small integers
    that e4m3 represents exactly (|piece| <= 16): square moduli r = sqrt(p) hi + lo with hi = rint(r / sqrt(p)); the others
    r = 16 hi + lo with hi = sign(r) ceil(|r| / 16), plus the Karatsuba piece hi + lo (mod.hpp:159-189).  The device evaluates this
    with binary32 magic-constant arithmetic (no conversion instructions); the same arithmetic is replayed here with numpy float32
    for ALL residues and checked against the integer definitions."""
    f32 = np.float32
    magic = f32(12582912.0)
    mods = T.moduli("FP8")
    for idx, p in enumerate(mods):
        h = p // 2
        r = np.arange(-h, h + 1, dtype=np.int64)
        af = (np.int32(0x4B400000) + r.astype(np.int32)).view(np.float32) - magic      # int -> float without I2F
        assert np.array_equal(af.astype(np.int64), r)
        if idx < 6:
            sq = f32(T.FP8_SQRT_MODULI[idx])
            assert int(sq) ** 2 == p
            inv = f32(1.0) / sq
            hf = ((af * inv) + magic) - magic
            lf = af - sq * hf                                                             # exact: small integers
            want_h = np.rint((r.astype(np.float32) * inv)).astype(np.int64)
            assert np.array_equal(hf.astype(np.int64), want_h)
            assert np.array_equal((int(sq) * hf.astype(np.int64) + lf.astype(np.int64)), r)
            pieces = [hf, lf]
        else:
            y = np.abs(af) * f32(0.0625) + f32(0.9375 - 0.46875)
            hm = (y + magic) - magic
            hf = np.copysign(hm, af)
            lf = af - f32(16.0) * hf
            want_h = np.sign(r) * -(-np.abs(r) // 16)                                    # sign(r) * ceil(|r| / 16)
            assert np.array_equal(hf.astype(np.int64), want_h), p
            assert np.array_equal(16 * hf.astype(np.int64) + lf.astype(np.int64), r)
            pieces = [hf, lf, hf + lf]
        for pc in pieces:
            v = pc.astype(np.int64)
            assert np.array_equal(v.astype(np.float32), pc) and np.abs(v).max() <= 16, (p, np.abs(v).max())
