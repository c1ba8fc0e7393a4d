"""CPU replays of the integer / bit-level formulas the CUDA kernels rely on (host logic only; no GPU, no oracle):
each is checked against the plain mathematical definition, exhaustively where the domain allows."""
import numpy as np
import pytest

from gemmul8_b200 import tables as T


def _sym_wrap(r, p):
    h = p >> 1
    r = np.where(r > h, r - p, r)
    return np.where(r < -h, r + p, r)


@pytest.mark.parametrize("be", ["INT8", "FP8"])
def test_epilogue_mod_i32(be):
    """g8_common.cuh:mod_i32 (GEMM epilogues, combine passes, residue sums): x - p * mulhi(x, floor(2^32 / p)) followed by ONE
    symmetric wrap is the symmetric residue for every |x| < 2^31 (p = 256 / 1024 keep +p/2, which the narrow store wraps)."""
    rng = np.random.default_rng(1)
    for p in T.moduli(be):
        pinv = (1 << 32) // p
        edge = np.array([0, 1, -1, 2 ** 31 - 1, -(2 ** 31 - 1), -(2 ** 31)], dtype=np.int64)
        near = (np.arange(-3, 4)[:, None] + (np.arange(-40, 41) * p)[None, :]).reshape(-1).astype(np.int64)
        big = (rng.integers(-(2 ** 31) // p, (2 ** 31) // p, size=2000) * p)[:, None] + np.arange(-2, 3)[None, :]
        x = np.concatenate([edge, near, big.reshape(-1), rng.integers(-(2 ** 31), 2 ** 31, size=400000)])
        x = x[(x >= -(2 ** 31)) & (x < 2 ** 31)]
        q = (x * pinv) >> 32                       # signed 32 x 32 -> high word (arithmetic shift of the exact 64-bit product)
        r = _sym_wrap(x - p * q, p)
        h = p >> 1
        assert np.all((r - x) % p == 0)
        assert r.min() >= -h and r.max() <= h
        if p % 2:                                   # odd modulus: the unique symmetric representative
            assert np.array_equal(r, ((x + h) % p) - h)
        else:                                       # even modulus: mulhi under-estimates the quotient, so the tie is always +p/2 --
            assert np.array_equal(r, h - ((h - x) % p))   # the result is a function of the residue CLASS (range (-p/2, p/2])


def test_crt_int8_to_double_splice():
    """g8_crt.cu: the residue byte c (int8) becomes a double without a conversion instruction: (c + 128) is spliced into the
    mantissa of 2^52 and (2^52 + 128) is subtracted -- exact for all 256 bytes; int16 likewise with 32768 (FP8 backend)."""
    c = np.arange(-128, 128, dtype=np.int64)
    bits = (np.uint64(0x43300000) << np.uint64(32)) | ((c & 0xFF) ^ 0x80).astype(np.uint64)
    assert np.array_equal(bits.view(np.float64) - (2.0 ** 52 + 128.0), c.astype(np.float64))
    c = np.arange(-32768, 32768, dtype=np.int64)
    bits = (np.uint64(0x43300000) << np.uint64(32)) | ((c & 0xFFFF) ^ 0x8000).astype(np.uint64)
    assert np.array_equal(bits.view(np.float64) - (2.0 ** 52 + 32768.0), c.astype(np.float64))


def test_residue_sum_by_dp4a_selectors():
    """K-sharded owner side: sum over shards of int8 residues with dp4a against one-hot selectors, then mod p: equals the residue of
    the exact sum (|sum| <= 8 * 127) for every modulus."""
    rng = np.random.default_rng(2)
    for p in T.moduli("INT8"):
        h = p // 2
        parts = rng.integers(-min(h, 127), min(h, 127) + 1, size=(8, 4096)).astype(np.int64)
        tot = parts.sum(axis=0)
        pinv = (1 << 32) // p
        r = _sym_wrap(tot - p * ((tot * pinv) >> 32), p)
        assert np.all((r - tot) % p == 0) and np.abs(r).max() <= h


def test_fp8_shard_sum_is_canonical():
    """K-sharded FP8 owner side (i16_sum_parts_kernel): every shard delivers mod_i32 of ITS partial sum; mod_i32 of the sum of those
    equals mod_i32 of the un-sharded sum for every modulus, ties of the even moduli included -- the owner's C_mid is the single-GPU one."""
    rng = np.random.default_rng(4)

    def mod_i32(x, p):
        return _sym_wrap(x - p * ((x * ((1 << 32) // p)) >> 32), p)

    for p in T.moduli("FP8"):
        h = p // 2
        parts = rng.integers(-(2 ** 24), 2 ** 24, size=(8, 20000)).astype(np.int64)
        parts[:, :64] = (np.arange(64) - 32) * h            # multiples of p / 2: the tie cases
        shard = mod_i32(parts, p)
        assert np.array_equal(mod_i32(shard.sum(axis=0), p), mod_i32(parts.sum(axis=0), p))
        assert np.abs(shard.sum(axis=0)).max() < 2 ** 15 * 8


def test_fp8_recombination_formulas():
    """mod.hpp:106-130 as used by f8_combine_kernel: with a = s hi_a + lo_a, b = s hi_b + lo_b (s^2 = p) the three products
    c0 = hi_a lo_b, c1 = lo_a hi_b, c2 = lo_a lo_b give a b = s (c0 + c1) + c2 (mod p); Karatsuba pieces (16 hi + lo, hi + lo)
    give a b = 256 c0 + 16 (c2 - c0 - c1) + c1 with c0 = hi hi, c1 = lo lo, c2 = (hi + lo)(hi + lo)."""
    rng = np.random.default_rng(3)
    mods = T.moduli("FP8")
    for idx, p in enumerate(mods):
        h = p // 2
        a = rng.integers(-h, h + 1, size=5000)
        b = rng.integers(-h, h + 1, size=5000)
        if idx < 6:
            s = T.FP8_SQRT_MODULI[idx]
            ha, hb = np.rint(a / s).astype(np.int64), np.rint(b / s).astype(np.int64)
            la, lb = a - s * ha, b - s * hb
            t = s * (ha * lb + la * hb) + la * lb
        else:
            ha, hb = np.sign(a) * -(-np.abs(a) // 16), np.sign(b) * -(-np.abs(b) // 16)
            la, lb = a - 16 * ha, b - 16 * hb
            c0, c1, c2 = ha * hb, la * lb, (ha + la) * (hb + lb)
            t = 256 * c0 + 16 * (c2 - c0 - c1) + c1
        assert np.all((t - a * b) % p == 0), p
