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
#   level 1 (binary64, once per group of three (INT8) / two (FP8) moduli with product M_g < 2^24):
#            t = fma(x, 1/M_g, 1.5 * 2^52); q = t - 1.5 * 2^52 = rint(x / M_g); r = fma(-M_g, q, x)   (|r| <= 0.52 M_g, exact integer)
#            a = int(r + B_g)  with the ONE group bias  B_g = (M_g - 1) / 2  [+ M_g for FP8]:  since M_g = 0 (mod p) for every member p,
#            B_g = (p - 1) / 2 = h (mod p) for ALL members at once, i.e. (a mod p) - h is the symmetric residue of x for each of them.
#            INT8 only: a < 0 (rounding slop of q) is lifted by + M_g:  a = umin(a, a + M_g).
#   level 2 (32-bit integer pipe, per modulus, TWO instructions; direct remainder by a multiply-high, Lemire-Kaser-Kurz 2019):
#            low = (a * ceil(2^32 / p)) mod 2^32;   s = hi32(low * p) + (-h)     = (a mod p) - h  in [-h, h]
#            exact iff a * e < 2^32 with e = ceil(2^32 / p) * p - 2^32 -- `fast_mod_tables` asserts it for the whole range of a
#            of every modulus, tests/test_tables.py replays the device arithmetic exhaustively.
#   power-of-two modulus (256 INT8 / 1024 FP8): x = r + M_0 q exactly, and the low word of t IS q mod 2^32, hence
#            x mod 2^k = (a_0 - B_0 + M_0 * lo32(t)) mod 2^k   -- two integer instructions, no extra floating-point reduction.
# ---------------------------------------------------------------------------------------------
MAGIC_RINT = 1.5 * 2.0 ** 52


def fast_mod_tables(backend: str = "INT8"):
    """Groups, group products / biases and per-modulus level-2 constants.  INT8: p_0 = 256 is recovered from group 0 and the
    other moduli form groups of three; FP8: p_1 = 1024 likewise and the others form groups of two (products stay below 2^24 so
    that the level-1 remainder is an exact small integer and the level-2 argument stays inside the multiply-high range)."""
    mods = moduli(backend)
    pow2_idx = 0 if backend == "INT8" else 1
    gsz = 3 if backend == "INT8" else 2
    lift = backend == "INT8"  # negative a lifted by + M_g (INT8); FP8 has room to add M_g unconditionally
    order = [i for i in range(len(mods)) if i != pow2_idx]
    groups = [order[j:j + gsz] for j in range(0, len(order), gsz)]
    grpM, grp_bias = [], []
    magic, half = [0] * len(mods), [0] * len(mods)
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
        assert 2 ** 13 < M < 2 ** 24 and M % 2 == 1
        grpM.append(M)
        B = (M - 1) // 2 + (0 if lift else M)
        grp_bias.append(B)
        # |level-1 remainder| <= (0.5 + eps) M with eps = |x / M| * 2^-52 <= 2^-6.6 (x < 2^63, M > 2^17.6): allow 0.52 M
        vmax = int(0.52 * M) + 2
        amax = vmax + B
        assert (lift and vmax - B < M) or (not lift and B - vmax >= 0)      # one + M lifts every negative a / a is never negative
        for i in members:
            pm = mods[i]
            assert M % pm == 0 and (B - pm // 2) % pm == 0                   # the group bias is = h (mod p) for every member
            mg = -(-(1 << 32) // pm)  # ceil(2^32 / p)
            e = mg * pm - (1 << 32)
            assert 0 < e <= pm and amax * e < (1 << 32) and amax < 2 ** 31, (pm, e, amax)
            magic[i], half[i] = mg, pm // 2
    members = [list(g) + [-1] * (gsz - len(g)) for g in groups]
    # number of groups needed for the first n moduli
    ngroups = [0] * (len(mods) + 1)
    for n in range(len(mods) + 1):
        ngroups[n] = sum(1 for g in groups if g[0] < n)
    return dict(groups=groups, members=members, grpM=grpM, grp_bias=grp_bias, lift=lift, magic=magic, half=half, ngroups=ngroups,
                fold=20 if backend == "INT8" else 30)


def hexf(x: float) -> str:
    return float(x).hex()
