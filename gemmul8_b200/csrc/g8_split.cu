// gemmul8_b200 -- stage 1 of the Ozaki-II pipeline: shift selection + split of op(A), op(B) into
// int8 residue planes (one HBM pass per operand, all moduli fused).
//
// Replaces (reference GEMMul8/src): scaling_fast_real.hpp:27-217, scaling_fast_complex.hpp:9-206,
// scaling_accu_real.hpp:23-375, scaling_accu_complex.hpp:15-398, scaling.hpp, mod.hpp:294-355, find_max.hpp.
//
// Two access shapes exist for an operand "rows x inner" (rows = m for A / n for B, inner = k):
//   row-contiguous : row r is contiguous along the inner index  (A with op T/C, B with op N)
//   row-strided    : rows are interleaved, element (r,l) at X[l*ld + r]      (A with op N, B with op T/C)
// The row-contiguous kernels use one 256-thread block per row; the row-strided kernels transpose
// 32 x 128 tiles through XOR-swizzled shared memory so that both the global loads (along r) and the
// plane stores (8 B per lane, 128 B per row segment, along l) are fully coalesced.
//
// Bit-parity notes.  The residues are the unique symmetric representatives, so any correct modular
// arithmetic matches the reference.  The SHIFTS however depend on (a) MUFU.LG2 via __log2f and the
// directed-rounding intrinsic sequence and (b), in fast mode, on the ORDER of the round-up sum of
// squares (find_max.hpp:258-341).  Both are reproduced here operation for operation: partial sums are
// formed over the same index classes (l mod 256 per thread, resp. l mod 32 per lane) and combined by the
// same shuffle tree.
#include "g8_internal.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <map>
#include <utility>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

namespace g8 {

// ------------------------------------------------------------------------------------------------
// element helpers
// ------------------------------------------------------------------------------------------------
template <typename T> struct Scalar;
template <> struct Scalar<float> { using U = float; static constexpr bool cplx = false; };
template <> struct Scalar<double> { using U = double; static constexpr bool cplx = false; };
template <> struct Scalar<float2> { using U = float; static constexpr bool cplx = true; };
template <> struct Scalar<double2> { using U = double; static constexpr bool cplx = true; };

__device__ __forceinline__ float  fma_ru_(float a, float b, float c) { return __fmaf_ru(a, b, c); }
__device__ __forceinline__ double fma_ru_(double a, double b, double c) { return __fma_ru(a, b, c); }
__device__ __forceinline__ float  add_ru_(float a, float b) { return __fadd_ru(a, b); }
__device__ __forceinline__ double add_ru_(double a, double b) { return __dadd_ru(a, b); }
__device__ __forceinline__ int ilogb0(double x) { return x == 0.0 ? 0 : ilogb(x); }
__device__ __forceinline__ int ilogb0(float x) { return x == 0.0f ? 0 : ilogbf(x); }

// |x| statistics of one element: running max and round-up sum of squares (Tsqr_add_ru, template_math.hpp:45-49)
template <typename T> __device__ __forceinline__ void acc_stats(const T &v, typename Scalar<T>::U &amax, typename Scalar<T>::U &sum) {
    if constexpr (Scalar<T>::cplx) {
        const auto x = fabs(v.x), y = fabs(v.y);
        amax         = max(max(x, y), amax);
        sum          = fma_ru_(y, y, fma_ru_(x, x, sum));
    } else {
        const auto x = fabs(v);
        amax         = max(x, amax);
        sum          = fma_ru_(x, x, sum);
    }
}
template <typename T> __device__ __forceinline__ void acc_amax(const T &v, typename Scalar<T>::U &amax) {
    if constexpr (Scalar<T>::cplx) amax = max(max(fabs(v.x), fabs(v.y)), amax);
    else amax = max(fabs(v), amax);
}

// shuffle-down tree, exactly the association order of the reference's inner_warp_sum/max (template_math.hpp:179-212)
template <typename U> __device__ __forceinline__ U warp_sum_ru(U s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = add_ru_(s, __shfl_down_sync(0xffffffffu, s, o));
    return s;
}
template <typename U> __device__ __forceinline__ U warp_max(U s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = max(s, __shfl_down_sync(0xffffffffu, s, o));
    return s;
}

// fast-mode shift (scaling_fast_real.hpp:6-22): same intrinsic sequence, runtime log2P
__device__ __forceinline__ int fast_shift(double amax, double vecnrm, float log2P) {
    const int exponent   = ilogb0(vecnrm);
    const float vecnrmf  = __double2float_ru(scalbn(vecnrm, -exponent));
    const float log2vsum = __fadd_ru(__log2f(vecnrmf), (float)exponent);
    const float log2vnrm = __fmul_ru(0x1.000006p-1f, log2vsum);
    const float exp1     = __fsub_rd(__fsub_rd(log2P, 1.5f), fmaxf(1.0f, log2vnrm));
    return __float2int_rd(exp1) - ilogb0((float)amax);
}
__device__ __forceinline__ int fast_shift(float amax, float vecnrm, float log2P) {
    const float log2vsum = __log2f(vecnrm);
    const float log2vnrm = __fmul_ru(0x1.000006p-1f, log2vsum);
    const float exp1     = __fsub_rd(__fsub_rd(log2P, 1.5f), fmaxf(1.0f, log2vnrm));
    return __float2int_rd(exp1) - ilogb0(amax);
}

// accurate mode, first shift: s0 = 5 - ilogb(amax)  (scaling_accu_real.hpp:39, maxUFP<INT8> = 5)
template <typename U> __device__ __forceinline__ int accu_s0(U amax) { return 5 - ilogb0(amax); }

// ceil(|a| * 2^s) -> int8, accurate-mode bound matrices.  Same case analysis as scaling.hpp:3-46
// (including its treatment of subnormal inputs), written on the raw bit fields.
template <typename U> __device__ __forceinline__ int8_t upper_bound_i8(U a, int s) {
    constexpr int prec = sizeof(U) == 8 ? 52 : 23;
    constexpr int bias = sizeof(U) == 8 ? 1023 : 127;
    constexpr int bits = sizeof(U) == 8 ? 64 : 32;
    uint64_t raw;
    if constexpr (sizeof(U) == 8) raw = (uint64_t)__double_as_longlong(a) & 0x7FFFFFFFFFFFFFFFull;
    else raw = (uint64_t)(__float_as_uint(a) & 0x7FFFFFFFu);
    if (raw == 0) return 0;
    const int bexp      = (int)(raw >> prec);
    const uint64_t frac = raw & ((1ull << prec) - 1);
    uint64_t mant;
    int e;
    if (bexp) {
        mant = frac | (1ull << prec);
        e    = bexp - bias;
    } else {
        const int k = (bits == 64 ? __clzll((long long)frac) : __clz((int)frac)) - (bits - prec);
        mant        = (frac << k) | (1ull << prec);
        e           = (1 - bias) - k;
    }
    e += s;
    const int shift = prec - e;
    if (shift <= 0) return (int8_t)(shift > -64 ? (mant << (-shift)) : 0);
    if (shift >= prec + 1) return (int8_t)1;
    const uint64_t mask = (1ull << shift) - 1;
    return (int8_t)((mant >> shift) + ((mant & mask) != 0));
}

// ------------------------------------------------------------------------------------------------
// Residues of the scaled integer x = trunc(a * 2^s) for all moduli.
//
// x is kept as a binary64 value (it is one: at most 53 significant bits).  Instead of the reference's
// per-modulus 64-bit mulhi chains (mod.hpp:31-55) we reduce in two levels (gemmul8_b200/tables.py:fast_mod_tables):
//   level 1 (FP64 pipe, once per group of three moduli, M_g < 2^24):
//       t = fma(x, 1/M_g, 1.5 2^52);  q = t - 1.5 2^52 = rint(x / M_g);  r = fma(-M_g, q, x);  a = int(r + B_g), B_g = (M_g - 1) / 2
//       B_g = (p - 1) / 2 (mod p) for EVERY member p of the group, so (a mod p) - h is the symmetric residue for all of them.
//   level 2 (integer pipe, per modulus, TWO instructions):  low = a * ceil(2^32 / p);  s = hi32(low * p) + (-h) = (a mod p) - h
//       (direct remainder by multiply-high; exact range asserted in tables.py, replayed exhaustively in tests/test_tables.py).
//   p = 256 (INT8) / 1024 (FP8): x = r + M_0 q exactly and lo32(t) = q mod 2^32, so x mod 2^k = (a_0 - B_0 + M_0 lo32(t)) mod 2^k.
// The results are the unique symmetric representatives, hence bit-identical to the reference's planes.
// LARGE (num_moduli > 15): |x| may exceed 2^63, so level 1 first folds x modulo M_g * 2^20 (a multiple of 2^k: the power-of-two
// residue is unchanged).
// ------------------------------------------------------------------------------------------------
constexpr double kMagic = 6755399441055744.0; // 1.5 * 2^52: (v + kMagic) - kMagic == rint(v) for |v| < 2^51

struct RowScale {
    double f1, f2; // 2^s split into two exactly representable factors (|s| may exceed 1023)
};
__device__ __forceinline__ RowScale make_scale(int s) {
    s = max(-2044, min(2046, s)); // shifts of finite inputs stay within +-1200; keep the exponent fields valid regardless
    const int s1 = s / 2, s2 = s - s1;
    RowScale r;
    r.f1 = __longlong_as_double((long long)(1023 + s1) << 52);
    r.f2 = __longlong_as_double((long long)(1023 + s2) << 52);
    return r;
}
// trunc(a * 2^s): both products are exact whenever the result is >= 1 in magnitude ("split")
template <typename U> __device__ __forceinline__ double scaled_trunc(U a, const RowScale &sc) {
    return trunc(__dmul_rn(__dmul_rn((double)a, sc.f1), sc.f2));
}
__device__ __forceinline__ double group_rem(double x, double M, double invM) {
    const double q = __dadd_rn(__fma_rn(x, invM, kMagic), -kMagic);
    return __fma_rn(-M, q, x);
}
// level 1 of group g: a = int(x mod M_g + B_g) >= 0.  POW2: also x mod 2^32-ish seed of the power-of-two modulus (only its low
// 8 / 10 bits are meaningful), from group 0.
template <bool LARGE, int BE, bool POW2> __device__ __forceinline__ int32_t level1(double x, int g, int32_t &pow2) {
    const double M = g8d_grpM[BE][g], invM = g8d_grpInvM[BE][g];
    if constexpr (LARGE) x = group_rem(x, M * g8d_foldMul[BE][0], invM * g8d_foldMul[BE][1]);
    const double t = __fma_rn(x, invM, kMagic); // low word = rint(x / M) mod 2^32
    const double r = __fma_rn(-M, __dadd_rn(t, -kMagic), x);
    int32_t a      = __double2loint(__dadd_rn(r, g8d_grpBias[BE][g]));
    if constexpr (POW2) pow2 = a + g8d_grpMint[BE][g] * __double2loint(t) - g8d_grpBiasInt[BE][g];
    // INT8: the rounding slop of q can leave a slightly negative; one + M_g lifts it (unsigned min picks the non-wrapped one)
    if constexpr (BE == INT8) a = (int32_t)min((uint32_t)a, (uint32_t)a + (uint32_t)g8d_grpMint[BE][g]);
    return a;
}
// symmetric residue (as int32; the low byte is the int8 plane value) of a (= x + h mod p) modulo moduli[idx]
template <int BE> __device__ __forceinline__ int32_t level2(int32_t a, int idx) {
    const uint32_t low = (uint32_t)a * g8d_mmagic[BE][idx];
    uint32_t s;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(low), "r"((uint32_t)g8d_moduli[BE][idx]), "r"(0u - (uint32_t)g8d_mhalf[BE][idx]));
    return (int32_t)s;
}
// bytes {a, b, c, d}.b0 -> one 32-bit word
__device__ __forceinline__ uint32_t pack4(int32_t a, int32_t b, int32_t c, int32_t d) {
    return __byte_perm(__byte_perm((uint32_t)a, (uint32_t)b, 0x0040), __byte_perm((uint32_t)c, (uint32_t)d, 0x0040), 0x5410);
}
// (Re + Im) mod p, symmetric, from the two symmetric residues (mod.hpp:327)
__device__ __forceinline__ int32_t add_wrap(int32_t r, int32_t q, int32_t p) {
    return sym_wrap((int)(int8_t)r + (int)(int8_t)q, p);
}

// One thread's NV consecutive inner indices of one row -> all planes.  REAL: x[NV]; planes[0].  CPLX: xr/xi; planes[0..2].
template <int NW> __device__ __forceinline__ void store_words(int8_t *dst, const uint32_t (&w)[NW]) {
    if constexpr (NW == 4) *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    else if constexpr (NW == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
    else *reinterpret_cast<uint32_t *>(dst) = w[0];
}
// one level-1 group (FIRST: group 0, which also yields the p = 256 plane) and the planes of its member moduli
template <bool LARGE, int NV, bool CPLX, bool FIRST>
__device__ __forceinline__ void split_group(const double (&xr)[NV], const double (&xi)[NV], int g, int num_moduli, int8_t *const (&planes)[3],
                                            size_t plane_stride, size_t off) {
    constexpr int NW = NV / 4;
    int32_t ar[NV], ai[NV];
    if constexpr (FIRST) {
        int32_t zr[NV], zi[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            ar[j] = level1<LARGE, INT8, true>(xr[j], g, zr[j]);
            if constexpr (CPLX) ai[j] = level1<LARGE, INT8, true>(xi[j], g, zi[j]);
        }
        uint32_t w0[NW], w1[NW], w2[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            w0[q] = pack4(zr[4 * q], zr[4 * q + 1], zr[4 * q + 2], zr[4 * q + 3]);
            if constexpr (CPLX) {
                w1[q] = pack4(zi[4 * q], zi[4 * q + 1], zi[4 * q + 2], zi[4 * q + 3]);
                w2[q] = pack4(add_wrap(zr[4 * q], zi[4 * q], 256), add_wrap(zr[4 * q + 1], zi[4 * q + 1], 256),
                              add_wrap(zr[4 * q + 2], zi[4 * q + 2], 256), add_wrap(zr[4 * q + 3], zi[4 * q + 3], 256));
            }
        }
        store_words<NW>(planes[0] + off, w0);
        if constexpr (CPLX) store_words<NW>(planes[1] + off, w1), store_words<NW>(planes[2] + off, w2);
    } else {
        int32_t unused;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            ar[j] = level1<LARGE, INT8, false>(xr[j], g, unused);
            if constexpr (CPLX) ai[j] = level1<LARGE, INT8, false>(xi[j], g, unused);
        }
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int idx = 3 * g + 1 + t;
        if (idx < num_moduli) {
            uint32_t w0[NW], w1[NW], w2[NW];
            const int32_t hp = g8d_mhalf[INT8][idx] + g8d_moduli[INT8][idx]; // (Re + Im): sr + si + h + p >= 0, = sr + si + h (mod p)
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                int32_t a[4], b[4], c[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    a[j] = level2<INT8>(ar[4 * q + j], idx);
                    if constexpr (CPLX) {
                        b[j] = level2<INT8>(ai[4 * q + j], idx);
                        c[j] = level2<INT8>(a[j] + b[j] + hp, idx);
                    }
                }
                w0[q] = pack4(a[0], a[1], a[2], a[3]);
                if constexpr (CPLX) w1[q] = pack4(b[0], b[1], b[2], b[3]), w2[q] = pack4(c[0], c[1], c[2], c[3]);
            }
            const size_t po = (size_t)idx * plane_stride + off;
            store_words<NW>(planes[0] + po, w0);
            if constexpr (CPLX) store_words<NW>(planes[1] + po, w1), store_words<NW>(planes[2] + po, w2);
        }
    }
}
template <bool LARGE, int NV, bool CPLX>
__device__ __forceinline__ void split_store(const double (&xr)[NV], const double (&xi)[NV], int num_moduli, int8_t *const (&planes)[3],
                                            size_t plane_stride, size_t off) {
    split_group<LARGE, NV, CPLX, true>(xr, xi, 0, num_moduli, planes, plane_stride, off);
    const int ngroups = g8d_numGroups[INT8][num_moduli];
    for (int g = 1; g < ngroups; ++g) split_group<LARGE, NV, CPLX, false>(xr, xi, g, num_moduli, planes, plane_stride, off);
}

// ------------------------------------------------------------------------------------------------
// FP8 backend (real types): every residue r (|r| <= p/2 <= 544) is written as 2 or 3 small integers that e4m3
// represents exactly (mod.hpp:159-189): square moduli p = s^2: r = s*hi + lo, hi = rintf(r/s); other moduli:
// r = 16*hi + lo with hi = sign(r)*ceil(|r|/16), plus the Karatsuba plane hi + lo.  Plane order as table.hpp:69-75.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fp8_of_int(int32_t v) { // |v| <= 16: exact in e4m3
    return (uint32_t)__nv_cvt_float_to_fp8((float)v, __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ uint32_t pack4u(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return a | (b << 8) | (c << 16) | (d << 24); }

// REAL: x -> planes[0].  CPLX: (xr, xi) -> Re planes[0], Im planes[1], (Re + Im) mod p planes[2] (mod.hpp:315-355 for the third set)
template <bool LARGE, int NV, bool CPLX>
__device__ __forceinline__ void split_store_f8(const double (&x)[NV], const double (&xi)[NV], int num_moduli, int8_t *const (&planes)[3],
                                               size_t plane_stride, size_t off) {
    constexpr int NW = NV / 4;
    auto emit = [&](int8_t *base, int idx, const int32_t (&r)[NV]) {
        auto store = [&](int plane, const uint32_t (&w)[NW]) {
            int8_t *dst = base + (size_t)plane * plane_stride + off;
            if constexpr (NW == 4) *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
            else if constexpr (NW == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
            else *reinterpret_cast<uint32_t *>(dst) = w[0];
        };
        const int pbase = idx < 6 ? 2 * idx : 12 + 3 * (idx - 6);
        // All piece arithmetic runs on the FMA pipe: int -> float by the 1.5 * 2^23 magic constant, rounding by adding / subtracting
        // it, and ONE packed F2FP per two pieces.  (I2F / F2I / FRND / single-value F2FP all issue on the quarter-rate conversion
        // pipe, which bounded this kernel.)  Every value involved is an integer or a multiple of 1/64 below 2^10: exact in binary32.
        constexpr float kMagicF = 12582912.0f; // 1.5 * 2^23
        float af[NV], hf[NV], lf[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) af[j] = __fsub_rn(__int_as_float(0x4B400000 + r[j]), kMagicF); // exact: |r| < 2^22
        if (idx < 6) {
            const float sq = (float)g8d_f8sqrt[idx], inv = __fdiv_rn(1.0f, sq);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                hf[j] = __fsub_rn(__fadd_rn(__fmul_rn(af[j], inv), kMagicF), kMagicF); // rintf(af * inv): round-to-nearest-even like rintf
                lf[j] = __fmaf_rn(-sq, hf[j], af[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                // h = sign(a) * ceil(|a| / 16) = sign(a) * floor((|a| + 15) / 16); (|a| + 15) / 16 is a multiple of 1/16, so
                // floor(y) = rint(y - 15/32)
                const float y  = __fmaf_rn(fabsf(af[j]), 0.0625f, 0.9375f - 0.46875f);
                const float hm = __fsub_rn(__fadd_rn(y, kMagicF), kMagicF);
                hf[j]          = copysignf(hm, af[j]);
                lf[j]          = __fmaf_rn(-16.0f, hf[j], af[j]);
            }
        }
        auto pack8 = [&](const float (&v)[NV], uint32_t (&w)[NW]) {
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                const uint32_t a = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(v[4 * q], v[4 * q + 1]), __NV_SATFINITE, __NV_E4M3);
                const uint32_t b = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(v[4 * q + 2], v[4 * q + 3]), __NV_SATFINITE, __NV_E4M3);
                w[q]             = a | (b << 16);
            }
        };
        uint32_t w[NW];
        pack8(hf, w);
        store(pbase, w);
        pack8(lf, w);
        store(pbase + 1, w);
        if (idx >= 6) {
            float sf[NV];
#pragma unroll
            for (int j = 0; j < NV; ++j) sf[j] = __fadd_rn(hf[j], lf[j]);
            pack8(sf, w);
            store(pbase + 2, w);
        }
    };
    // one modulus: residues of the real (and imaginary, and their sum) parts -> plane sets
    auto emit_all = [&](int idx, const int32_t (&rr)[NV], const int32_t (&ri)[NV]) {
        emit(planes[0], idx, rr);
        if constexpr (CPLX) {
            emit(planes[1], idx, ri);
            const int32_t p = g8d_moduli[FP8][idx];
            int32_t rs[NV];
#pragma unroll
            for (int j = 0; j < NV; ++j) rs[j] = sym_wrap(rr[j] + ri[j], p);
            emit(planes[2], idx, rs);
        }
    };
    // group 0 = moduli {0, 2}; it also yields modulus index 1 (p = 1024) from the low bits of x (see level1)
    const int ngroups = g8d_numGroups[FP8][num_moduli];
    for (int g = 0; g < ngroups; ++g) {
        int32_t ar[NV], ai[NV];
        if (g == 0) {
            int32_t r[NV], ri[NV];
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                int32_t z;
                ar[j] = level1<LARGE, FP8, true>(x[j], 0, z);
                z &= 1023;
                r[j] = z > 512 ? z - 1024 : z; // symmetric, +512 kept (mod.hpp:79-93)
                if constexpr (CPLX) {
                    ai[j] = level1<LARGE, FP8, true>(xi[j], 0, z);
                    z &= 1023;
                    ri[j] = z > 512 ? z - 1024 : z;
                } else {
                    ai[j] = 0, ri[j] = 0;
                }
            }
            if (num_moduli > 1) emit_all(1, r, ri);
        } else {
            int32_t unused;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                ar[j] = level1<LARGE, FP8, false>(x[j], g, unused);
                if constexpr (CPLX) ai[j] = level1<LARGE, FP8, false>(xi[j], g, unused);
                else ai[j] = 0;
            }
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int idx = g8d_grpMembers[FP8][g][t];
            if (idx >= 0 && idx < num_moduli) {
                int32_t r[NV], ri[NV];
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    r[j] = level2<FP8>(ar[j], idx);
                    if constexpr (CPLX) ri[j] = level2<FP8>(ai[j], idx);
                    else ri[j] = 0;
                }
                emit_all(idx, r, ri);
            }
        }
    }
}

// accurate mode, FP8: round-up conversion of |a| * 2^s0 (< 2^8) to e4m3 (scaling.hpp:48-54,80-85)
template <typename U> __device__ __forceinline__ uint32_t upper_bound_f8(U a, const double f1, const double f2) {
    const double v = __dmul_rn(__dmul_rn(fabs((double)a), f1), f2); // exact scaling
    uint32_t r;
    if constexpr (sizeof(U) == 8) r = (uint32_t)__nv_cvt_double_to_fp8(v, __NV_SATFINITE, __NV_E4M3);
    else r = (uint32_t)__nv_cvt_float_to_fp8((float)v, __NV_SATFINITE, __NV_E4M3);
    const float back = __half2float(__half(__nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)r, __NV_E4M3)));
    return r + (uint32_t)((double)back < v);
}

// NV consecutive inner indices of one row: either the accurate-mode bound plane(s) (MODE 2, sft = s0) or all residue planes
template <typename T, bool LARGE, int MODE, int NV, int BE>
__device__ __forceinline__ void emit_elements(const T (&v)[NV], int sft, const SplitArgs &a, size_t off) {
    using U             = typename Scalar<T>::U;
    constexpr bool CPLX = Scalar<T>::cplx;
    constexpr int NW    = NV / 4;
    if constexpr (BE == FP8) {
        const RowScale scale = make_scale(sft);
        auto st = [&](int8_t *dst, const uint32_t (&w)[NW]) {
            if constexpr (NW == 4) *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
            else if constexpr (NW == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
            else *reinterpret_cast<uint32_t *>(dst) = w[0];
        };
        if constexpr (MODE == 2) {
            // bound planes: |x| * 2^s0 rounded UP to e4m3; complex: |Re| -> planes[0], |Im| -> planes[1]
            uint32_t w[NW], wi[NW];
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                if constexpr (CPLX) {
                    w[q]  = pack4u(upper_bound_f8<U>(v[4 * q].x, scale.f1, scale.f2), upper_bound_f8<U>(v[4 * q + 1].x, scale.f1, scale.f2),
                                   upper_bound_f8<U>(v[4 * q + 2].x, scale.f1, scale.f2), upper_bound_f8<U>(v[4 * q + 3].x, scale.f1, scale.f2));
                    wi[q] = pack4u(upper_bound_f8<U>(v[4 * q].y, scale.f1, scale.f2), upper_bound_f8<U>(v[4 * q + 1].y, scale.f1, scale.f2),
                                   upper_bound_f8<U>(v[4 * q + 2].y, scale.f1, scale.f2), upper_bound_f8<U>(v[4 * q + 3].y, scale.f1, scale.f2));
                } else {
                    w[q] = pack4u(upper_bound_f8<U>(v[4 * q], scale.f1, scale.f2), upper_bound_f8<U>(v[4 * q + 1], scale.f1, scale.f2),
                                  upper_bound_f8<U>(v[4 * q + 2], scale.f1, scale.f2), upper_bound_f8<U>(v[4 * q + 3], scale.f1, scale.f2));
                }
            }
            st(a.planes[0] + off, w);
            if constexpr (CPLX) st(a.planes[1] + off, wi);
        } else {
            double x[NV], xi[NV];
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                if constexpr (CPLX) x[j] = scaled_trunc<U>(v[j].x, scale), xi[j] = scaled_trunc<U>(v[j].y, scale);
                else x[j] = scaled_trunc<U>(v[j], scale), xi[j] = 0.0;
            }
            split_store_f8<LARGE, NV, CPLX>(x, xi, a.num_moduli, a.planes, a.plane_stride, off);
        }
    } else if constexpr (MODE == 2) {
        uint32_t wr[NW], wi[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            if constexpr (CPLX) {
                wr[q] = pack4(upper_bound_i8<U>(v[4 * q].x, sft), upper_bound_i8<U>(v[4 * q + 1].x, sft),
                              upper_bound_i8<U>(v[4 * q + 2].x, sft), upper_bound_i8<U>(v[4 * q + 3].x, sft));
                wi[q] = pack4(upper_bound_i8<U>(v[4 * q].y, sft), upper_bound_i8<U>(v[4 * q + 1].y, sft),
                              upper_bound_i8<U>(v[4 * q + 2].y, sft), upper_bound_i8<U>(v[4 * q + 3].y, sft));
            } else {
                wr[q] = pack4(upper_bound_i8<U>(v[4 * q], sft), upper_bound_i8<U>(v[4 * q + 1], sft),
                              upper_bound_i8<U>(v[4 * q + 2], sft), upper_bound_i8<U>(v[4 * q + 3], sft));
            }
        }
        auto st = [&](int8_t *dst, const uint32_t (&w)[NW]) {
            if constexpr (NW == 4) *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
            else if constexpr (NW == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
            else *reinterpret_cast<uint32_t *>(dst) = w[0];
        };
        st(a.planes[0] + off, wr);
        if constexpr (CPLX) st(a.planes[1] + off, wi);
    } else {
        const RowScale scale = make_scale(sft);
        double xr[NV], xi[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            if constexpr (CPLX) xr[j] = scaled_trunc<U>(v[j].x, scale), xi[j] = scaled_trunc<U>(v[j].y, scale);
            else xr[j] = scaled_trunc<U>(v[j], scale), xi[j] = 0.0;
        }
        split_store<LARGE, NV, CPLX>(xr, xi, a.num_moduli, a.planes, a.plane_stride, off);
    }
}

// ------------------------------------------------------------------------------------------------
// ROW-CONTIGUOUS kernels: one block (256 threads) per row of op(X)
// MODE 0: split with the stored shift; MODE 1: fast mode (stats -> shift -> split);
// MODE 2: accurate stage (i): amax -> s0, write bound plane(s)
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T ldg_conj(const T *p, bool conj) {
    T v = __ldg(p);
    if constexpr (Scalar<T>::cplx)
        if (conj) v.y = -v.y;
    return v;
}

template <typename T, bool LARGE, int MODE, int BE>
__global__ void __launch_bounds__(256) split_rowcontig_kernel(SplitArgs a) {
    using U       = typename Scalar<T>::U;
    const T *in   = reinterpret_cast<const T *>(a.X) + (size_t)blockIdx.x * a.ld;
    const int row = blockIdx.x;
    const int k   = (int)a.inner;
    __shared__ U s_max[32], s_sum[32];
    int sft;

    if constexpr (MODE == 0) {
        sft = -(int)a.sft[row];
    } else if constexpr (MODE == 3) {
        sft = (int)a.sft[row]; // bound planes with an externally supplied s0 (K-sharded runs: s0 comes from the GLOBAL amax)
    } else {
        // thread t visits l = t, t+256, ... in order (find_max.hpp:272-277 / 26-38)
        U amax = 0, sum = 0;
#pragma unroll 4
        for (int l = threadIdx.x; l < k; l += 256) {
            const T v = __ldg(in + l);
            if constexpr (MODE == 1) acc_stats<T>(v, amax, sum);
            else acc_amax<T>(v, amax);
        }
        amax = warp_max(amax);
        if constexpr (MODE == 1) sum = warp_sum_ru(sum);
        if ((threadIdx.x & 31) == 0) {
            s_max[threadIdx.x >> 5] = amax;
            if constexpr (MODE == 1) s_sum[threadIdx.x >> 5] = sum;
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            sum = 0;
            if (threadIdx.x < 8) {
                amax = s_max[threadIdx.x];
                if constexpr (MODE == 1) sum = s_sum[threadIdx.x];
            }
            amax = warp_max(amax);
            if constexpr (MODE == 1) sum = warp_sum_ru(sum);
            if (threadIdx.x == 0) {
                s_max[0] = amax;
                if constexpr (MODE == 1) s_sum[0] = sum;
            }
        }
        __syncthreads();
        amax = s_max[0];
        if constexpr (MODE == 1) {
            sft = fast_shift(amax, s_sum[0], g8d_log2P[BE][a.num_moduli]);
            if (threadIdx.x == 0) a.sft[row] = (int16_t)(-sft);
        } else {
            sft = accu_s0(amax) + (BE == FP8 ? 2 : 0); // maxUFP: 5 (INT8) / 7 (FP8), template_type.hpp:147
            if (threadIdx.x == 0) a.sft[row] = (int16_t)sft;
        }
    }

    // 8 consecutive inner indices per thread -> one 64-bit store per plane, 256 B per warp.  (The second read of the row -- after the
    // statistics pass of modes 1 / 2 -- mostly hits L2; parking the row in shared memory instead was measured SLOWER, r02d sweep:
    // the 64 KB per block cut the occupancy to three blocks per SM.)
    constexpr int NV     = 8;
    const size_t row_off = (size_t)row * a.k_pad;
    for (int l = threadIdx.x * NV; l < (int)a.k_pad; l += 256 * NV) {
        T v[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            if (l + j < k) v[j] = ldg_conj(in + l + j, a.conj);
            else v[j] = T{};
        }
        emit_elements<T, LARGE, (MODE == 3 ? 2 : MODE), NV, BE>(v, sft, a, row_off + l);
    }
}

// ------------------------------------------------------------------------------------------------
// ROW-STRIDED statistics: block (32 lanes = rows, 32 partial classes); thread (x, y) walks l = y, y+32, ...
// of row blockIdx.x*32 + x, then lane-transposes through shared memory and reduces over the 32 classes
// (find_max.hpp:40-64, 306-341).  MODE 1: fast shift, MODE 2: accurate s0.
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE, int BE>
__global__ void __launch_bounds__(1024) stats_rowstrided_kernel(SplitArgs a) {
    using U = typename Scalar<T>::U;
    __shared__ U s_max[32][33], s_sum[32][33];
    const T *X   = reinterpret_cast<const T *>(a.X);
    const int x  = threadIdx.x, y = threadIdx.y;
    int row      = blockIdx.x * 32 + x;
    U amax = 0, sum = 0;
    if (row < (int)a.rows) {
        const T *rp = X + row;
#pragma unroll 4
        for (int l = y; l < (int)a.inner; l += 32) {
            const T v = __ldg(rp + (size_t)l * a.ld);
            if constexpr (MODE == 1) acc_stats<T>(v, amax, sum);
            else acc_amax<T>(v, amax);
        }
    }
    s_max[y][x] = amax;
    if constexpr (MODE == 1) s_sum[y][x] = sum;
    __syncthreads();
    amax = warp_max(s_max[x][y]);
    if constexpr (MODE == 1) sum = warp_sum_ru(s_sum[x][y]);
    row = blockIdx.x * 32 + y;
    if (row < (int)a.rows && x == 0) {
        if constexpr (MODE == 1) a.sft[row] = (int16_t)(-fast_shift(amax, sum, g8d_log2P[BE][a.num_moduli]));
        else a.sft[row] = (int16_t)(accu_s0(amax) + (BE == FP8 ? 2 : 0));
    }
}

// ------------------------------------------------------------------------------------------------
// ROW-STRIDED split / extract: tile = 32 rows x TL inner, 256 threads.
//   load   : lane = row (coalesced along r), 8 warps stride over l; raw values -> smem[l][r ^ swz(l)]
//   compute: thread (rr = t/8 [+32-row tile], seg = t%8) owns 16 consecutive l of one row, emits one
//            16-byte store per plane; 8 lanes cover a 128-byte line.
// MODE 0: residues of trunc(x * 2^-sft); MODE 2: bound plane(s) with s0 = sft (as stored)
// ------------------------------------------------------------------------------------------------
constexpr int RS_TL = 128; // row-strided tile: 32 rows x 128 inner; 16 segments of 8 per row
// one tile = 32 rows x 128 inner; lane = row (coalesced along r), the 16 warps stride over l: every thread moves 8 elements.
// Split into "global -> registers" and "registers -> XOR-swizzled shared memory" so that the loads of the NEXT tile can be in
// flight while the current one is reduced (register double buffering).
template <typename T> __device__ __forceinline__ void rowstrided_fetch_tile(const SplitArgs &a, T (&v)[RS_TL / 16], int r0, int l0) {
    const T *X     = reinterpret_cast<const T *>(a.X);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5; // 16 warps
    const int r    = r0 + lane;
#pragma unroll
    for (int j = 0; j < RS_TL / 16; ++j) {
        const int l = l0 + warp + 16 * j;
        v[j]        = T{};
        if (r < (int)a.rows && l < (int)a.inner) v[j] = ldg_conj(X + (size_t)l * a.ld + r, a.conj);
    }
}
template <typename T> __device__ __forceinline__ void rowstrided_park_tile(const T (&v)[RS_TL / 16], T *tile) {
    constexpr int SW = (sizeof(T) == 16) ? 0 : 1; // swizzle granularity that keeps both phases bank-conflict free
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < RS_TL / 16; ++j) {
        const int ll = warp + 16 * j;
        tile[ll * 32 + (lane ^ ((ll >> 3) << SW))] = v[j];
    }
}
template <typename T> __device__ __forceinline__ void rowstrided_load_tile(const SplitArgs &a, T *tile, int r0, int l0) {
    T v[RS_TL / 16];
    rowstrided_fetch_tile<T>(a, v, r0, l0);
    rowstrided_park_tile<T>(v, tile);
}
// thread (rr = t / 16, seg = t % 16) owns 8 consecutive l of one row of the tile and emits one 8-byte store per plane
template <typename T, bool LARGE, int MODE, int BE>
__device__ __forceinline__ void rowstrided_emit_tile(const SplitArgs &a, const T *tile, int r0, int l0, int sft) {
    constexpr int NV = 8;
    constexpr int SW = (sizeof(T) == 16) ? 0 : 1;
    const int rr  = threadIdx.x >> 4; // 0..31
    const int seg = threadIdx.x & 15; // 8 inner indices each
    const int row = r0 + rr;
    if (row >= (int)a.rows) return;
    const size_t off = (size_t)row * a.k_pad + l0 + seg * NV;
    T v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int ll = seg * NV + j;
        v[j]         = tile[ll * 32 + (rr ^ (seg << SW))];
    }
    emit_elements<T, LARGE, MODE, NV, BE>(v, sft, a, off);
}

// blockIdx.y owns `tiles_per_block` consecutive tiles along l; the loads of tile t + 1 are issued before tile t is reduced.
template <typename T, bool LARGE, int MODE, int BE>
__global__ void __launch_bounds__(512) split_rowstrided_kernel(SplitArgs a, int tiles_per_block) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *tile = reinterpret_cast<T *>(smem_raw); // [TL][32], column index XOR-swizzled by the segment number
    const int r0 = blockIdx.x * 32;
    const int ntiles = (int)(a.k_pad / RS_TL);
    const int t_begin = blockIdx.y * tiles_per_block, t_end = min(ntiles, t_begin + tiles_per_block);
    const int row = r0 + (threadIdx.x >> 4);
    int sft = 0;
    if (row < (int)a.rows) sft = (MODE == 0) ? -(int)a.sft[row] : (int)a.sft[row];
    T v[RS_TL / 16];
    rowstrided_fetch_tile<T>(a, v, r0, t_begin * RS_TL);
    for (int t = t_begin; t < t_end; ++t) {
        rowstrided_park_tile<T>(v, tile);
        __syncthreads();
        if (t + 1 < t_end) rowstrided_fetch_tile<T>(a, v, r0, (t + 1) * RS_TL); // in flight during the reduction below
        rowstrided_emit_tile<T, LARGE, MODE, BE>(a, tile, r0, t * RS_TL, sft);
        __syncthreads();
    }
}

// accurate mode, stage (iii): sft = -(s0 + floor(fmaf_rd(-0x1.000006p-1f, log2f(float(max)), log2P)))
// (scaling_accu_real.hpp:6-11,157-159,202-204); max[] was filled by the bound GEMM's atomicMax epilogue.
__global__ void finalize_accu_shift_kernel(int16_t *__restrict__ sft, const int32_t *__restrict__ cmax, int count, int num_moduli, int backend) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    // INT8: integer maxima; FP8: the maxima are non-negative floats stored by their bit pattern (scaling_accu_real.hpp:13-18)
    const float log2amax = __log2f(backend == INT8 ? __int2float_rn(cmax[i]) : __int_as_float(cmax[i]));
    const int g          = __float2int_rd(__fmaf_rd(-0x1.000006p-1f, log2amax, g8d_log2P[backend][num_moduli]));
    int s                = sft[i];
    s += g;
    sft[i] = (int16_t)(-s);
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
// |x| can exceed 2^63 (two level-1 folds needed) once num_moduli passes the backend's M threshold (common.hpp:15-27)
static bool is_large(int backend, int num_moduli) { return num_moduli > thresholds(backend).M; }

// dynamic shared memory beyond 48 KB needs an opt-in per kernel AND per device: done once per (kernel, device, size high-water mark),
// never repeated on the launch path.  The attribute is set to exactly what is requested (static + dynamic must stay <= 227 KB).
static void ensure_smem(const void *kern, size_t smem) {
    if (smem <= 48 * 1024) return;
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> granted;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    size_t &have = granted[{kern, dev}];
    if (smem > have && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) have = smem;
}
static int env_flag(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <typename T, int MODE, int BE> static void launch_rowcontig(const SplitArgs &a, bool large, cudaStream_t st) {
    const dim3 grid((unsigned)a.rows);
    if (MODE >= 2 || !large) split_rowcontig_kernel<T, false, MODE, BE><<<grid, 256, 0, st>>>(a);
    else split_rowcontig_kernel<T, true, MODE, BE><<<grid, 256, 0, st>>>(a);
}

template <typename T, int MODE, int BE> static void launch_rowstrided(const SplitArgs &a, bool large, cudaStream_t st) {
    // several tiles per block (software-pipelined), but still >= ~8 blocks per SM for load balance
    static const int tpb_pref = env_flag("G8_SPLIT_TILES_PER_BLOCK", 4);
    const int ntiles = (int)(a.k_pad / RS_TL);
    const size_t row_blocks = (a.rows + 31) / 32;
    int tpb = std::max(1, std::min(tpb_pref, ntiles));
    while (tpb > 1 && row_blocks * ((ntiles + tpb - 1) / tpb) < 1200) tpb >>= 1;
    const dim3 grid((unsigned)row_blocks, (unsigned)((ntiles + tpb - 1) / tpb));
    const size_t smem = RS_TL * 32 * sizeof(T);
    auto go = [&](auto kern) {
        ensure_smem(reinterpret_cast<const void *>(kern), smem);
        kern<<<grid, 512, smem, st>>>(a, tpb);
    };
    if (MODE == 2 || !large) go(split_rowstrided_kernel<T, false, MODE, BE>);
    else go(split_rowstrided_kernel<T, true, MODE, BE>);
}

template <typename T, int BE> static void split_typed(const SplitArgs &a, int mode, cudaStream_t st) {
    const bool large = is_large(BE, a.num_moduli);
    if (a.row_contig) {
        if (mode == 0) launch_rowcontig<T, 0, BE>(a, large, st);
        else if (mode == 1) launch_rowcontig<T, 1, BE>(a, large, st);
        else if (mode == 2) launch_rowcontig<T, 2, BE>(a, large, st);
        else launch_rowcontig<T, 3, BE>(a, large, st);
    } else {
        const dim3 sgrid((unsigned)((a.rows + 31) / 32)), sblock(32, 32);
        if (mode == 1) stats_rowstrided_kernel<T, 1, BE><<<sgrid, sblock, 0, st>>>(a);
        if (mode == 2) stats_rowstrided_kernel<T, 2, BE><<<sgrid, sblock, 0, st>>>(a);
        if (mode >= 2) launch_rowstrided<T, 2, BE>(a, large, st);
        else launch_rowstrided<T, 0, BE>(a, large, st);
    }
}

// mode 0: split with stored shifts, 1: fast (shift + split), 2: accurate stage (i) (s0 + bound planes),
// 3: bound planes with the stored s0 (no statistics pass)
void launch_split(const SplitArgs &a, int dtype, int mode, cudaStream_t st) {
    if (a.backend == FP8) {
        switch (dtype) {
        case F32: split_typed<float, FP8>(a, mode, st); break;
        case F64: split_typed<double, FP8>(a, mode, st); break;
        case C32: split_typed<float2, FP8>(a, mode, st); break;
        case C64: split_typed<double2, FP8>(a, mode, st); break;
        }
        return;
    }
    switch (dtype) {
    case F32: split_typed<float, INT8>(a, mode, st); break;
    case F64: split_typed<double, INT8>(a, mode, st); break;
    case C32: split_typed<float2, INT8>(a, mode, st); break;
    case C64: split_typed<double2, INT8>(a, mode, st); break;
    }
}

// ------------------------------------------------------------------------------------------------
// K-sharded multi-GPU support: local row statistics as doubles, and shifts from (globally reduced) statistics.
// There is no reference counterpart (the reference is single-GPU); the shift formulas are the reference's.
// ------------------------------------------------------------------------------------------------
template <typename T> __global__ void __launch_bounds__(256) stats_only_rowcontig_kernel(SplitArgs a, double *amax_out, double *sumsq_out) {
    using U     = typename Scalar<T>::U;
    const T *in = reinterpret_cast<const T *>(a.X) + (size_t)blockIdx.x * a.ld;
    __shared__ double s_max[8], s_sum[8];
    U amax = 0;
    double sum = 0;
    for (int l = threadIdx.x; l < (int)a.inner; l += 256) {
        const T v = __ldg(in + l);
        acc_amax<T>(v, amax);
        if constexpr (Scalar<T>::cplx) sum = __fma_ru((double)v.y, (double)v.y, __fma_ru((double)v.x, (double)v.x, sum));
        else sum = __fma_ru((double)v, (double)v, sum);
    }
    double dmax = warp_max((double)amax);
    sum         = warp_sum_ru(sum);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = dmax, s_sum[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) dmax = max(dmax, s_max[w]), sum = __dadd_ru(sum, s_sum[w]);
        amax_out[blockIdx.x] = dmax, sumsq_out[blockIdx.x] = sum;
    }
}
template <typename T> __global__ void __launch_bounds__(1024) stats_only_rowstrided_kernel(SplitArgs a, double *amax_out, double *sumsq_out) {
    using U = typename Scalar<T>::U;
    __shared__ double s_max[32][33], s_sum[32][33];
    const T *X  = reinterpret_cast<const T *>(a.X);
    const int x = threadIdx.x, y = threadIdx.y;
    int row     = blockIdx.x * 32 + x;
    U amax = 0;
    double sum = 0;
    if (row < (int)a.rows)
        for (int l = y; l < (int)a.inner; l += 32) {
            const T v = __ldg(X + row + (size_t)l * a.ld);
            acc_amax<T>(v, amax);
            if constexpr (Scalar<T>::cplx) sum = __fma_ru((double)v.y, (double)v.y, __fma_ru((double)v.x, (double)v.x, sum));
            else sum = __fma_ru((double)v, (double)v, sum);
        }
    s_max[y][x] = (double)amax, s_sum[y][x] = sum;
    __syncthreads();
    const double dmax = warp_max(s_max[x][y]);
    sum               = warp_sum_ru(s_sum[x][y]);
    row               = blockIdx.x * 32 + y;
    if (row < (int)a.rows && x == 0) amax_out[row] = dmax, sumsq_out[row] = sum;
}
// kind 0: fast-mode shift from (amax, sumsq); kind 1: accurate-mode s0 from amax
__global__ void shift_from_stats_kernel(const double *amax, const double *sumsq, int count, int num_moduli, int kind, int16_t *sft, int backend) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (kind == 0) {
        // the cross-rank sum was rounded to nearest: nudge it up so the Cauchy-Schwarz bound stays rigorous
        const double s = sumsq[i] * (1.0 + 0x1p-48);
        sft[i] = (int16_t)(-fast_shift(amax[i], s, g8d_log2P[backend][num_moduli]));
    } else {
        sft[i] = (int16_t)(accu_s0(amax[i]) + (backend == FP8 ? 2 : 0)); // maxUFP: 5 (INT8) / 7 (FP8), template_type.hpp:147
    }
}

void launch_stats(const SplitArgs &a, int dtype, double *amax, double *sumsq, cudaStream_t st) {
    const dim3 sgrid((unsigned)((a.rows + 31) / 32)), sblock(32, 32);
#define G8_STATS(T)                                                                                         \
    if (a.row_contig) stats_only_rowcontig_kernel<T><<<(unsigned)a.rows, 256, 0, st>>>(a, amax, sumsq);    \
    else stats_only_rowstrided_kernel<T><<<sgrid, sblock, 0, st>>>(a, amax, sumsq);
    switch (dtype) {
    case F32: G8_STATS(float) break;
    case F64: G8_STATS(double) break;
    case C32: G8_STATS(float2) break;
    case C64: G8_STATS(double2) break;
    }
#undef G8_STATS
}
void launch_shift_from_stats(const double *amax, const double *sumsq, size_t count, int num_moduli, int kind, int16_t *sft, cudaStream_t st, int backend) {
    shift_from_stats_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(amax, sumsq, (int)count, num_moduli, kind, sft, backend);
}

void launch_finalize_accu_shift(int16_t *sft, const int32_t *cmax, size_t count, int num_moduli, cudaStream_t st, int backend) {
    finalize_accu_shift_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(sft, cmax, (int)count, num_moduli, backend);
}

} // namespace g8
