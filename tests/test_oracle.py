"""The CPU oracle (oracle/g8_oracle.c + oracle/oracle.py) is pinned against
  * the reference's only known-answer vector (sample/dgemm_cuBLASLt_int8.cu:26-40 -> tests/golden/sample_kat.json),
  * an independent big-integer / exact-rational model written here (Appendix A of SURVEY.md),
  * size-independent properties (CRT value == exact integer product, symmetric residues)."""
import json
from fractions import Fraction
from pathlib import Path

import numpy as np
import pytest

from gemmul8_b200 import tables as T
from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent


def kat():
    g = json.loads((ROOT / "tests/golden/sample_kat.json").read_text())
    A = np.array([float.fromhex(x) for x in g["A"]]).reshape(5, 4).T
    B = np.array([float.fromhex(x) for x in g["B"]]).reshape(3, 5).T
    C = np.array([float.fromhex(x) for x in g["C_exact"]]).reshape(3, 4).T
    return A, B, C, g


def test_kat_exact_at_reference_settings():
    A, B, Cx, g = kat()
    r = O.emulate(A, B, num_moduli=g["num_moduli"], fastmode=g["fastmode"])
    assert np.array_equal(r["C"], Cx)  # N=15, accurate mode reproduces the exact product bit for bit
    assert not r["ambA"].any() and not r["ambB"].any()


@pytest.mark.parametrize("N,tol", [(14, 2e-15), (18, 3e-15), (20, 3e-15), (8, 1e-7)])
def test_kat_other_moduli_counts(N, tol):
    A, B, Cx, _ = kat()
    for fast in (False, True):
        r = O.emulate(A, B, num_moduli=N, fastmode=fast)
        assert np.linalg.norm(r["C"] - Cx) < tol * (30 if fast else 1)


def test_kat_exact_product_is_correctly_rounded():
    A, B, Cx, _ = kat()
    for i in range(4):
        for j in range(3):
            s = sum(Fraction(A[i, l]) * Fraction(B[l, j]) for l in range(5))
            assert float(s) == Cx[i, j]


# ---------------------------------------------------------------- independent exact model
def fma_exact(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def model(A, B, sftA, sftB, N, dd):
    m, k = A.shape
    n = B.shape[1]
    mods = T.moduli("INT8")[:N]
    trunc = lambda x, s: int(Fraction(x) * Fraction(2) ** s) if x >= 0 else -int(Fraction(-x) * Fraction(2) ** s)
    Ai = [[trunc(A[i, l], -int(sftA[i])) for l in range(k)] for i in range(m)]
    Bi = [[trunc(B[l, j], -int(sftB[j])) for l in range(k)] for j in range(n)]
    C = np.zeros((m, n))
    exact = np.zeros((m, n), dtype=object)
    for i in range(m):
        for j in range(n):
            res = []
            for p in mods:
                h = sum(T.sym_mod(a, p) * T.sym_mod(b, p) for a, b in zip(Ai[i], Bi[j]))
                c = T.sym_mod(h, p)
                res.append(c if not (p == 256 and c == 128) else -128)  # int8 store of +128
            Px, Py = T.P_dd("INT8", N)
            invP = T.invP("INT8", N)
            if dd:
                hi = lo = 0.0
                for (wx, wy), c in zip(T.qPi_2("INT8", N), res):
                    hi = fma_exact(wx, float(c), hi)
                    lo = fma_exact(wy, float(c), lo)
                q = float(round(Fraction(invP * hi)))  # rint: ties-to-even never hit here
                v = fma_exact(Py, q, fma_exact(Px, q, hi) + lo)
            else:
                s = 0.0
                for w, c in zip(T.qPi_1("INT8", N), res):
                    s = fma_exact(w, float(c), s)
                q = float(round(Fraction(invP * s)))
                v = fma_exact(Px, q, s)
            C[i, j] = float(Fraction(v) * Fraction(2) ** (int(sftA[i]) + int(sftB[j])))
            exact[i, j] = sum(a * b for a, b in zip(Ai[i], Bi[j]))
    return C, exact


@pytest.mark.parametrize("N", [3, 6, 7, 14, 16, 20])
def test_oracle_matches_independent_model(N):
    rng = np.random.default_rng(N)
    m, n, k = 5, 4, 9
    A = (rng.random((m, k)) - 0.5) * np.exp(rng.standard_normal((m, k)))
    B = (rng.random((k, n)) - 0.5) * np.exp(rng.standard_normal((k, n)))
    r = O.emulate(A, B, num_moduli=N, fastmode=False)
    Cm, exact = model(A, B, r["sftA"], r["sftB"], N, dd=N > 6)
    assert np.array_equal(r["C"], Cm)
    # CRT reconstructs the exact integer product A'B' (|A'B'| < P/2), up to the final double rounding
    for i in range(m):
        for j in range(n):
            want = float(Fraction(int(exact[i, j])) * Fraction(2) ** (int(r["sftA"][i]) + int(r["sftB"][j])))
            if N > 6:  # hi/lo chain: the hi sums are error-free by construction (table.hpp:329-331)
                assert abs(r["C"][i, j] - want) <= 2 * np.spacing(abs(want)) + 1e-300
            else:      # single chain: rounding errors of the N FMAs, each below ulp(P * 2^8)
                scale = 2.0 ** (int(r["sftA"][i]) + int(r["sftB"][j]))
                assert abs(r["C"][i, j] - want) <= N * T.prod("INT8", N) * 2.0 ** (8 - 52) * scale
            assert abs(int(exact[i, j])) * 2 < T.prod("INT8", N)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("fast", [False, True])
def test_oracle_accuracy_all_types(dtype, fast):
    rng = np.random.default_rng(5)
    m, n, k = 17, 13, 40
    cplx = np.dtype(dtype).kind == "c"
    def mk(s):
        x = (rng.random(s) - 0.5) * np.exp(rng.standard_normal(s) * 0.5)
        return (x + 1j * (rng.random(s) - 0.5)).astype(dtype) if cplx else x.astype(dtype)
    A, B = mk((m, k)), mk((k, n))
    N = 6 if np.dtype(dtype).itemsize in (4, 8) and np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else 15
    for opA, opB in (("N", "N"), ("T", "C")):
        As = A if opA == "N" else A.T.copy()
        Bs = B if opB == "N" else B.conj().T.copy()
        r = O.emulate(As, Bs, opA, opB, N, fast)
        ref = A.astype(np.complex128 if cplx else np.float64) @ B.astype(np.complex128 if cplx else np.float64)
        tol = 3e-5 if N == 6 else 1e-12
        assert np.abs(r["C"] - ref).max() <= tol * np.abs(ref).max()
        # residues are symmetric representatives
        for planes in r["A_lo"] + r["B_lo"]:
            for i, p in enumerate(T.moduli("INT8")[:N]):
                lim = 128 if p == 256 else p // 2
                assert planes[i].min() >= -lim and planes[i].max() <= (127 if p == 256 else lim)


def test_alpha_beta_modes():
    rng = np.random.default_rng(9)
    A, B = rng.standard_normal((6, 7)), rng.standard_normal((7, 5))
    C0 = rng.standard_normal((6, 5))
    base = O.emulate(A, B, num_moduli=14)["C"]
    for a, b in ((1, 1), (-1, 0), (-1, 1), (0.5, 2.0)):
        r = O.emulate(A, B, num_moduli=14, alpha=a, beta=b, C0=C0)["C"]
        assert np.allclose(r, a * base + b * C0, rtol=1e-15, atol=1e-15)


@pytest.mark.parametrize("dtype,N", [(np.complex128, 13), (np.complex64, 6), (np.float64, 12)])
def test_oracle_fp8_backend_accuracy(dtype, N):
    """FP8-backend restatement (16-bit residues, real and complex): fast mode reproduces the product to the emulated precision"""
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    m, n, k = 12, 9, 40
    def rnd(s):
        x = rng.standard_normal(s)
        return (x + 1j * rng.standard_normal(s)).astype(dtype) if np.dtype(dtype).kind == "c" else x.astype(dtype)
    A, B = rnd((m, k)), rnd((k, n))
    r = O.emulate(A, B, "N", "N", N, True, backend="FP8")
    wide = np.complex128 if np.dtype(dtype).kind == "c" else np.float64
    ref = A.astype(wide) @ B.astype(wide)
    err = np.abs(r["C"] - ref).max() / np.abs(ref).max()
    assert err < 64 * np.finfo(np.dtype(dtype)).eps, err


def _mod_i32(x, p):
    """csrc/g8_common.cuh:mod_i32 replayed on int64 arrays (see tests/test_device_formulas.py)"""
    r = x - p * ((x * ((1 << 32) // p)) >> 32)
    h = p >> 1
    r = np.where(r > h, r - p, r)
    return np.where(r < -h, r + p, r)


@pytest.mark.parametrize("dtype,backend,N", [(np.float64, "INT8", 9), (np.float64, "FP8", 8), (np.complex128, "FP8", 7), (np.complex64, "INT8", 5)])
def test_k_sharded_residues_sum_to_the_unsharded_ones(dtype, backend, N):
    """The algebra behind the K-sharded multi-GPU paths (csrc/g8_mg.cu), at the oracle level: with the GLOBAL shifts, every shard's
    C_mid (the residues of its partial product) summed over the shards and reduced again with the device's mod_i32 equals the C_mid of
    the un-sharded call bit for bit -- INT8 (crt_parts / i8_cplx_combine_parts) and FP8 (i16_sum_parts), real and complex."""
    rng = np.random.default_rng(77)
    m, n, kl, W = 24, 20, 40, 3
    cplx = np.dtype(dtype).kind == "c"
    def rand(shape):
        x = rng.standard_normal(shape) * np.exp(rng.standard_normal(shape))
        return (x + 1j * rng.standard_normal(shape)).astype(dtype) if cplx else x.astype(dtype)
    A, B = rand((m, kl * W)), rand((kl * W, n))
    full = O.emulate(A, B, num_moduli=N, fastmode=False, backend=backend)
    tot = None
    for s in range(W):
        ks = slice(s * kl, (s + 1) * kl)
        part = O.emulate(A[:, ks], B[ks, :], num_moduli=N, fastmode=False, backend=backend, sftA=full["sftA"], sftB=full["sftB"])
        tot = part["C_mid"].astype(np.int64) if tot is None else tot + part["C_mid"].astype(np.int64)
    mods = T.moduli(backend)[:N]
    for i, p in enumerate(mods):
        got = _mod_i32(tot[i], p)
        want = full["C_mid"][i].astype(np.int64)
        if backend == "INT8" and p == 256:
            got = got.astype(np.int8).astype(np.int64)     # +128 wraps in the int8 store, as in the reference
        assert np.array_equal(got, want), f"modulus {p}"
