"""Ozaki-II constants, derived from first principles with exact integer arithmetic.

The reference ships these numbers as literal tables (GEMMul8/src/table.hpp:9-849).  They are
mathematical facts of the moduli set, so we *derive* them instead of copying them:

  moduli      table.hpp:12-53   pairwise-coprime p_i (INT8: <=256, FP8: <=1089)
  P, invP     table.hpp:78-151  P_N = prod_{i<N} p_i, stored NEGATED as a double-double; 1/P_N
  log2P       table.hpp:161-203 RD_float( log2(P_N - 1)/2 - 0.5 )
  mod_pow2    table.hpp:209-258 symmetric residues of 2^j mod p_i (for the N>=16 split path)
  qPi_1       table.hpp:277-327 double( q_i * P_N/p_i ),  q_i = (P_N/p_i)^-1 mod p_i
  qPi_2       table.hpp:329-550 hi/lo split of q_i*P_N/p_i so that the hi FMA chain is exact

`tests/test_tables.py` pins every derived value against the numbers parsed out of the reference's
table.hpp (tests/golden/ref_tables.json, made by tools/extract_ref_tables.py).

This module is host-side product code (the generator of csrc/g8_tables.h and the source of the
Python-visible constants); the oracle imports it too.
"""
from __future__ import annotations

import math
import struct
from fractions import Fraction

INT8_MODULI = (256, 255, 253, 251, 247, 241, 239, 233, 229, 227,
               223, 217, 211, 199, 197, 193, 191, 181, 179, 173)
FP8_MODULI = (1089, 1024, 961, 841, 625, 529, 511, 509, 503, 499,
              491, 487, 481, 479, 467, 463, 461, 457, 449, 443)
FP8_SQRT_MODULI = (33, 32, 31, 29, 25, 23)
NOT_KARATSUBA = 6

# common.hpp:15-27
THRESHOLD = {
    "INT8": dict(P_is_double=6, S=7, M=15, L=25),
    "FP8": dict(P_is_double=5, S=5, M=12, L=20),
}
MAX_MODULI = 20


def moduli(backend: str):
    return INT8_MODULI if backend == "INT8" else FP8_MODULI


def num_mat(backend: str, n: int) -> int:
    """Number of low-precision planes per operand (table.hpp:69-75)."""
    if backend == "INT8":
        return n
    return 2 * n if n <= NOT_KARATSUBA else 2 * NOT_KARATSUBA + 3 * (n - NOT_KARATSUBA)


def prod(backend: str, n: int) -> int:
    out = 1
    for p in moduli(backend)[:n]:
        out *= p
    return out


def _rn(x) -> float:
    """Correctly rounded (nearest-even) double of an int or Fraction."""
    if isinstance(x, int):
        return float(Fraction(x))  # Fraction -> float is a correctly rounded true division
    return float(x)


def P_dd(backend: str, n: int):
    """(-P) as an unevaluated double-double (x, y), x = RN(-P), y = RN(-P - x)."""
    P = prod(backend, n)
    hi = _rn(-P)
    lo = _rn(-P - int(hi))
    return hi, lo


def invP(backend: str, n: int) -> float:
    return _rn(Fraction(1, prod(backend, n)))


def _f32_round_down(x) -> float:
    """Largest binary32 value <= x (x an mpmath mpf or Fraction)."""
    import mpmath

    xf = float(x)
    f = struct.unpack("f", struct.pack("f", xf))[0]  # RN to float32
    if mpmath.mpf(f) > x:
        bits = struct.unpack("I", struct.pack("f", f))[0]
        bits = bits - 1 if f > 0 else bits + 1
        f = struct.unpack("f", struct.pack("I", bits))[0]
    return f


def log2P(backend: str, n: int) -> float:
    """RD_float( log2(P-1)/2 - 0.5 )  (table.hpp:159)."""
    import mpmath

    if n == 2:
        # The reference's literal for N=2 (table.hpp:164,185) sits 47 / 2 float ulps ABOVE the formula
        # (all 36 other entries match it exactly).  The shift exponents -- hence the output bits --
        # depend on this value, so bit parity requires the reference's number here.
        return float.fromhex("0x1.dfd1ecp+2" if backend == "INT8" else "0x1.316baep+3")
    with mpmath.workprec(400):
        v = mpmath.log(prod(backend, n) - 1, 2) / 2 - mpmath.mpf("0.5")
        return _f32_round_down(v)


def crt_weights(backend: str, n: int):
    """Exact integers w_i = q_i * P/p_i with w_i == 1 (mod p_i), w_i == 0 (mod p_j)."""
    P = prod(backend, n)
    out = []
    for p in moduli(backend)[:n]:
        Pi = P // p
        q = pow(Pi % p, -1, p)
        out.append(q * Pi)
    return out


def qPi_1(backend: str, n: int):
    return [_rn(w) for w in crt_weights(backend, n)]


def _hi_bits(backend: str, n: int) -> int:
    """Bits kept in the exact 'hi' part: 53 - ceil(log2(rho)), rho = sum_i floor(p_i/2) (table.hpp:329-331)."""
    rho = sum(p // 2 for p in moduli(backend)[:n])
    return 53 - math.ceil(math.log2(rho))


def qPi_2(backend: str, n: int):
    """[(hi, lo)]: every hi is w_i truncated to a COMMON quantum 2^e, e = bitlen(max_i w_i) - keep,
    so that sum_i hi_i*c_i (|c_i| <= floor(p_i/2)) is exact in binary64; lo = RN(w_i - hi)."""
    keep = _hi_bits(backend, n)
    ws = crt_weights(backend, n)
    drop = max(max(w.bit_length() for w in ws) - keep, 0)
    out = []
    for w in ws:
        hi_int = (w >> drop) << drop
        out.append((float(hi_int), _rn(w - hi_int)))
    return out


def sym_mod(a: int, p: int) -> int:
    """Symmetric residue as the reference's `wrapping` produces it (mod.hpp:8-12):
    value in [-floor(p/2), floor(p/2)]; for even p the +p/2 representative is kept."""
    r = a % p
    if r > p // 2:
        r -= p
    return r


def mod_pow2(backend: str):
    """Rows of symmetric 2^j mod p used when the scaled operand exceeds 63 bits (table.hpp:209-269)."""
    rows = []
    if backend == "INT8":
        for p in INT8_MODULI[1:]:
            rows.append([sym_mod(1 << (j + 7), p) for j in range(57)])
    else:
        for idx, p in enumerate(FP8_MODULI):
            if idx == 1:
                continue  # 1024: handled by bit masking
            rows.append([sym_mod(1 << (j + 8), p) for j in range(64)])
    return rows


# ---------------------------------------------------------------------------------------------
# Tables of the split kernel's two-level modular reduction (our own scheme, no reference counterpart):
#   level 1: r = x - M_g * rint(x / M_g) in binary64 for groups of three consecutive moduli (product M_g < 2^24),
#   level 2: per modulus p of the group, with the integer v = r (|v| <= 0.501 M_g):
#            a1 = v + h + K p  (>= 0),  q = floor(a1 / p) = umulhi(a1, ceil(2^32/p)),  s = (a1 - h) - q p  in [-h, h].
# `fast_mod_tables` proves the exact-floor condition a1 * (ceil(2^32/p) p - 2^32) < 2^32 for every modulus.
# ---------------------------------------------------------------------------------------------
MAGIC_RINT = 1.5 * 2.0 ** 52


def fast_mod_tables(backend: str = "INT8"):
    """Groups, group products and per-modulus level-2 constants.  INT8: p_0 = 256 is handled by byte extraction and the
    other moduli form groups of three; FP8: p_1 = 1024 is handled by masking and the others form groups of two (products
    must stay below 2^24 so that the level-1 remainder is an exact small integer)."""
    mods = moduli(backend)
    pow2_idx = 0 if backend == "INT8" else 1
    gsz = 3 if backend == "INT8" else 2
    order = [i for i in range(len(mods)) if i != pow2_idx]
    groups = [order[j:j + gsz] for j in range(0, len(order), gsz)]
    grpM, magic, half, bias = [], [0] * len(mods), [0] * len(mods), [0.0] * len(mods)
    for gi, members in enumerate(groups):
        # a short last group borrows preceding moduli for its product so that x / M_g stays below 2^51
        span = list(members)
        j = order.index(members[0])
        while len(span) < gsz:
            j -= 1
            span.insert(0, order[j])
        M = 1
        for i in span:
            M *= mods[i]
        assert 2 ** 13 < M < 2 ** 24
        grpM.append(M)
        # |level-1 remainder| <= (0.5 + eps) M with eps = |x / M| * 2^-52 <= 2^-6.6 (x < 2^63, M > 2^17.6): allow 0.52 M
        vmax = int(0.52 * M) + 2
        for i in members:
            pm = mods[i]
            h = pm // 2
            K = -(-(vmax + h) // pm) + 1
            mg = -(-(1 << 32) // pm)  # ceil(2^32 / p)
            e = mg * pm - (1 << 32)
            a1max = vmax + h + K * pm
            assert 0 < e <= pm and a1max * e < (1 << 32) and a1max < 2 ** 31, (pm, e, a1max)
            magic[i], half[i], bias[i] = mg, h, MAGIC_RINT + h + K * pm
    members = [list(g) + [-1] * (gsz - len(g)) for g in groups]
    # number of groups needed for the first n moduli
    ngroups = [0] * (len(mods) + 1)
    for n in range(len(mods) + 1):
        ngroups[n] = sum(1 for g in groups if g[0] < n)
    return dict(groups=groups, members=members, grpM=grpM, magic=magic, half=half, bias=bias, ngroups=ngroups,
                fold=20 if backend == "INT8" else 30)


def hexf(x: float) -> str:
    return float(x).hex()
