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
// plane stores (16 B per lane, 128 B per row segment, along l) are fully coalesced.
//
// Bit-parity notes.  The residues are the unique symmetric representatives, so any correct modular
// arithmetic matches the reference.  The SHIFTS however depend on (a) MUFU.LG2 via __log2f and the
// directed-rounding intrinsic sequence and (b), in fast mode, on the ORDER of the round-up sum of
// squares (find_max.hpp:258-341).  Both are reproduced here operation for operation: partial sums are
// formed over the same index classes (l mod 256 per thread, resp. l mod 32 per lane) and combined by the
// same shuffle tree.
#include "g8_internal.cuh"

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
// scaled integer A' = trunc(a * 2^s) kept as (sign, 53-bit magnitude, left shift) and its residues.
// REGIME 0: |A'| < 2^31 (N <= S), 1: |A'| < 2^63 (N <= M), 2: larger, shift > 10 possible (N > M).
// ------------------------------------------------------------------------------------------------
struct Scaled {
    int64_t v; // REGIME 0/1: the full signed value.  REGIME 2: signed 53-bit mantissa
    int sh;    // REGIME 2 only: remaining left shift (>= 0)
};

template <int REGIME, typename U> __device__ __forceinline__ Scaled scale_trunc(U a_in, int s) {
    const double a      = (double)a_in; // exact for float
    const uint64_t raw  = (uint64_t)__double_as_longlong(a);
    const int bexp      = (int)((raw >> 52) & 0x7FF);
    const uint64_t frac = raw & 0xFFFFFFFFFFFFFull;
    const uint64_t mant = bexp ? (frac | (1ull << 52)) : frac;
    const int sh        = (bexp ? bexp - 1075 : -1074) + s;
    Scaled r;
    r.sh = 0;
    uint64_t mag;
    if (sh <= 0) {
        mag = (sh > -64) ? (mant >> (-sh)) : 0ull;
    } else if (REGIME < 2 || sh <= 10) {
        mag = mant << (sh & 63);
    } else {
        mag  = mant;
        r.sh = sh;
    }
    r.v = (raw >> 63) ? -(int64_t)mag : (int64_t)mag;
    return r;
}

template <int REGIME> __device__ __forceinline__ int32_t residue(const Scaled &x, int idx, int32_t p) {
    if constexpr (REGIME == 0) {
        return mod_i32((int32_t)x.v, p, g8d_pinv32[INT8][idx]);
    } else if constexpr (REGIME == 1) {
        return mod_i64(x.v, p, g8d_pinv64[INT8][idx]);
    } else {
        const int32_t r = mod_i64(x.v, p, g8d_pinv64[INT8][idx]);
        if (x.sh == 0) return r;
        if (idx == 0) return 0; // p = 256 and shift > 10: multiple of 256
        const int32_t w = g8d_modpow2[INT8][idx - 1][x.sh - 7];
        return mod_i32(r * w, p, g8d_pinv32[INT8][idx]);
    }
}

__device__ __forceinline__ uint32_t pack4(int32_t a, int32_t b, int32_t c, int32_t d) {
    return (uint32_t)(a & 0xFF) | ((uint32_t)(b & 0xFF) << 8) | ((uint32_t)(c & 0xFF) << 16) | ((uint32_t)d << 24);
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

template <typename T, int REGIME, int MODE>
__global__ void __launch_bounds__(256) split_rowcontig_kernel(SplitArgs a) {
    using U       = typename Scalar<T>::U;
    const T *in   = reinterpret_cast<const T *>(a.X) + (size_t)blockIdx.x * a.ld;
    const int row = blockIdx.x;
    const int k   = (int)a.inner;
    __shared__ U s_max[32], s_sum[32];
    int sft;

    if constexpr (MODE == 0) {
        sft = -(int)a.sft[row];
    } else {
        // thread t visits l = t, t+256, ... in order (find_max.hpp:272-277 / 26-38)
        U amax = 0, sum = 0;
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
            sft = fast_shift(amax, s_sum[0], g8d_log2P[INT8][a.num_moduli]);
            if (threadIdx.x == 0) a.sft[row] = (int16_t)(-sft);
        } else {
            sft = accu_s0(amax);
            if (threadIdx.x == 0) a.sft[row] = (int16_t)sft;
        }
    }

    // 4 consecutive inner indices per thread -> one 32-bit store per plane, 128 B per warp
    const size_t row_off = (size_t)row * a.k_pad;
    for (int l4 = threadIdx.x; l4 < (int)(a.k_pad >> 2); l4 += 256) {
        const int l = l4 << 2;
        T v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (l + j < k) v[j] = ldg_conj(in + l + j, a.conj);
            else v[j] = T{};
        }
        if constexpr (MODE == 2) {
            if constexpr (Scalar<T>::cplx) {
                int32_t re[4], im[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    re[j] = upper_bound_i8<U>(v[j].x, sft);
                    im[j] = upper_bound_i8<U>(v[j].y, sft);
                }
                uint32_t *o0 = reinterpret_cast<uint32_t *>(a.planes[0] + row_off + l);
                uint32_t *o1 = reinterpret_cast<uint32_t *>(a.planes[1] + row_off + l);
                *o0          = pack4(re[0], re[1], re[2], re[3]);
                *o1          = pack4(im[0], im[1], im[2], im[3]);
            } else {
                uint32_t *o = reinterpret_cast<uint32_t *>(a.planes[0] + row_off + l);
                *o = pack4(upper_bound_i8<U>(v[0], sft), upper_bound_i8<U>(v[1], sft), upper_bound_i8<U>(v[2], sft),
                           upper_bound_i8<U>(v[3], sft));
            }
        } else {
            if constexpr (Scalar<T>::cplx) {
                Scaled xr[4], xi[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    xr[j] = scale_trunc<REGIME, U>(v[j].x, sft);
                    xi[j] = scale_trunc<REGIME, U>(v[j].y, sft);
                }
                for (int i = 0; i < a.num_moduli; ++i) {
                    const int32_t p = g8d_moduli[INT8][i];
                    int32_t r[4], q[4], s[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        r[j] = residue<REGIME>(xr[j], i, p);
                        q[j] = residue<REGIME>(xi[j], i, p);
                        s[j] = sym_wrap((int)(int8_t)r[j] + (int)(int8_t)q[j], p); // (Re+Im) mod p (mod.hpp:327)
                    }
                    const size_t off = (size_t)i * a.plane_stride + row_off + l;
                    *reinterpret_cast<uint32_t *>(a.planes[0] + off) = pack4(r[0], r[1], r[2], r[3]);
                    *reinterpret_cast<uint32_t *>(a.planes[1] + off) = pack4(q[0], q[1], q[2], q[3]);
                    *reinterpret_cast<uint32_t *>(a.planes[2] + off) = pack4(s[0], s[1], s[2], s[3]);
                }
            } else {
                Scaled x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = scale_trunc<REGIME, U>(v[j], sft);
                for (int i = 0; i < a.num_moduli; ++i) {
                    const int32_t p = g8d_moduli[INT8][i];
                    *reinterpret_cast<uint32_t *>(a.planes[0] + (size_t)i * a.plane_stride + row_off + l) =
                        pack4(residue<REGIME>(x[0], i, p), residue<REGIME>(x[1], i, p), residue<REGIME>(x[2], i, p),
                              residue<REGIME>(x[3], i, p));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ROW-STRIDED statistics: block (32 lanes = rows, 32 partial classes); thread (x, y) walks l = y, y+32, ...
// of row blockIdx.x*32 + x, then lane-transposes through shared memory and reduces over the 32 classes
// (find_max.hpp:40-64, 306-341).  MODE 1: fast shift, MODE 2: accurate s0.
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE>
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
        if constexpr (MODE == 1) a.sft[row] = (int16_t)(-fast_shift(amax, sum, g8d_log2P[INT8][a.num_moduli]));
        else a.sft[row] = (int16_t)accu_s0(amax);
    }
}

// ------------------------------------------------------------------------------------------------
// ROW-STRIDED split / extract: tile = 32 rows x TL inner, 256 threads.
//   load   : lane = row (coalesced along r), 8 warps stride over l; raw values -> smem[l][r ^ swz(l)]
//   compute: thread (rr = t/8 [+32-row tile], seg = t%8) owns 16 consecutive l of one row, emits one
//            16-byte store per plane; 8 lanes cover a 128-byte line.
// MODE 0: residues of trunc(x * 2^-sft); MODE 2: bound plane(s) with s0 = sft (as stored)
// ------------------------------------------------------------------------------------------------
template <typename T, int REGIME, int MODE>
__global__ void __launch_bounds__(256) split_rowstrided_kernel(SplitArgs a) {
    using U                = typename Scalar<T>::U;
    constexpr bool CPLX    = Scalar<T>::cplx;
    constexpr int TL       = 128;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *tile = reinterpret_cast<T *>(smem_raw); // [TL][32]

    const T *X     = reinterpret_cast<const T *>(a.X);
    const int r0   = blockIdx.x * 32;
    const int l0   = blockIdx.y * TL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    {
        const int r = r0 + lane;
#pragma unroll 4
        for (int j = 0; j < TL / 8; ++j) {
            const int ll = warp + 8 * j;
            const int l  = l0 + ll;
            T v{};
            if (r < (int)a.rows && l < (int)a.inner) v = ldg_conj(X + (size_t)l * a.ld + r, a.conj);
            tile[ll * 32 + (lane ^ ((ll >> 4) << 2))] = v;
        }
    }
    __syncthreads();

    const int rr  = threadIdx.x >> 3; // 0..31
    const int seg = threadIdx.x & 7;  // 16 inner indices each
    const int row = r0 + rr;
    if (row >= (int)a.rows) return;
    const int sft = (MODE == 0) ? -(int)a.sft[row] : (int)a.sft[row];
    const size_t off = (size_t)row * a.k_pad + l0 + seg * 16;
    if (l0 + seg * 16 >= (int)a.k_pad) return;

    T v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int ll = seg * 16 + j;
        v[j]         = tile[ll * 32 + (rr ^ ((ll >> 4) << 2))];
    }

    if constexpr (MODE == 2) {
        if constexpr (CPLX) {
            uint32_t wr[4], wi[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                wr[q] = pack4(upper_bound_i8<U>(v[4 * q].x, sft), upper_bound_i8<U>(v[4 * q + 1].x, sft),
                              upper_bound_i8<U>(v[4 * q + 2].x, sft), upper_bound_i8<U>(v[4 * q + 3].x, sft));
                wi[q] = pack4(upper_bound_i8<U>(v[4 * q].y, sft), upper_bound_i8<U>(v[4 * q + 1].y, sft),
                              upper_bound_i8<U>(v[4 * q + 2].y, sft), upper_bound_i8<U>(v[4 * q + 3].y, sft));
            }
            *reinterpret_cast<uint4 *>(a.planes[0] + off) = make_uint4(wr[0], wr[1], wr[2], wr[3]);
            *reinterpret_cast<uint4 *>(a.planes[1] + off) = make_uint4(wi[0], wi[1], wi[2], wi[3]);
        } else {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                w[q] = pack4(upper_bound_i8<U>(v[4 * q], sft), upper_bound_i8<U>(v[4 * q + 1], sft),
                             upper_bound_i8<U>(v[4 * q + 2], sft), upper_bound_i8<U>(v[4 * q + 3], sft));
            *reinterpret_cast<uint4 *>(a.planes[0] + off) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else {
        if constexpr (CPLX) {
            Scaled xr[16], xi[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                xr[j] = scale_trunc<REGIME, U>(v[j].x, sft);
                xi[j] = scale_trunc<REGIME, U>(v[j].y, sft);
            }
            for (int i = 0; i < a.num_moduli; ++i) {
                const int32_t p = g8d_moduli[INT8][i];
                uint32_t w0[4], w1[4], w2[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int32_t r[4], s[4], t[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        r[j] = residue<REGIME>(xr[4 * q + j], i, p);
                        s[j] = residue<REGIME>(xi[4 * q + j], i, p);
                        t[j] = sym_wrap((int)(int8_t)r[j] + (int)(int8_t)s[j], p);
                    }
                    w0[q] = pack4(r[0], r[1], r[2], r[3]);
                    w1[q] = pack4(s[0], s[1], s[2], s[3]);
                    w2[q] = pack4(t[0], t[1], t[2], t[3]);
                }
                const size_t o = (size_t)i * a.plane_stride + off;
                *reinterpret_cast<uint4 *>(a.planes[0] + o) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
                *reinterpret_cast<uint4 *>(a.planes[1] + o) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
                *reinterpret_cast<uint4 *>(a.planes[2] + o) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
            }
        } else {
            Scaled x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = scale_trunc<REGIME, U>(v[j], sft);
            for (int i = 0; i < a.num_moduli; ++i) {
                const int32_t p = g8d_moduli[INT8][i];
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    w[q] = pack4(residue<REGIME>(x[4 * q], i, p), residue<REGIME>(x[4 * q + 1], i, p),
                                 residue<REGIME>(x[4 * q + 2], i, p), residue<REGIME>(x[4 * q + 3], i, p));
                *reinterpret_cast<uint4 *>(a.planes[0] + (size_t)i * a.plane_stride + off) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
}

// complex accurate mode: third bound plane = Re - Im of the two bound planes (scaling_accu_complex.hpp:5,46)
__global__ void bound_diff_kernel(const int8_t *__restrict__ re, const int8_t *__restrict__ im, int8_t *__restrict__ out, size_t n16) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    const uint4 a = reinterpret_cast<const uint4 *>(re)[i], b = reinterpret_cast<const uint4 *>(im)[i];
    uint4 o;
    o.x = __vsub4(a.x, b.x);
    o.y = __vsub4(a.y, b.y);
    o.z = __vsub4(a.z, b.z);
    o.w = __vsub4(a.w, b.w);
    reinterpret_cast<uint4 *>(out)[i] = o;
}

// accurate mode, stage (iii): sft = -(s0 + floor(fmaf_rd(-0x1.000006p-1f, log2f(float(max)), log2P)))
// (scaling_accu_real.hpp:6-11,157-159,202-204); max[] was filled by the bound GEMM's atomicMax epilogue.
__global__ void finalize_accu_shift_kernel(int16_t *__restrict__ sft, const int32_t *__restrict__ cmax, int count, int num_moduli) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float log2amax = __log2f(__int2float_rn(cmax[i]));
    const int g          = __float2int_rd(__fmaf_rd(-0x1.000006p-1f, log2amax, g8d_log2P[INT8][num_moduli]));
    int s                = sft[i];
    s += g;
    sft[i] = (int16_t)(-s);
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
static int regime_of(int num_moduli) {
    const Thresholds t = thresholds(INT8);
    return num_moduli <= t.S ? 0 : (num_moduli <= t.M ? 1 : 2);
}

template <typename T, int MODE> static void launch_rowcontig(const SplitArgs &a, int regime, cudaStream_t st) {
    const dim3 grid((unsigned)a.rows);
    if (MODE == 2 || regime == 0) split_rowcontig_kernel<T, 0, MODE><<<grid, 256, 0, st>>>(a);
    else if (regime == 1) split_rowcontig_kernel<T, 1, MODE><<<grid, 256, 0, st>>>(a);
    else split_rowcontig_kernel<T, 2, MODE><<<grid, 256, 0, st>>>(a);
}

template <typename T, int MODE> static void launch_rowstrided(const SplitArgs &a, int regime, cudaStream_t st) {
    const dim3 grid((unsigned)((a.rows + 31) / 32), (unsigned)(a.k_pad / 128));
    const size_t smem = 128 * 32 * sizeof(T);
    auto go = [&](auto kern) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, 256, smem, st>>>(a);
    };
    if (MODE == 2 || regime == 0) go(split_rowstrided_kernel<T, 0, MODE>);
    else if (regime == 1) go(split_rowstrided_kernel<T, 1, MODE>);
    else go(split_rowstrided_kernel<T, 2, MODE>);
}

template <typename T> static void split_typed(const SplitArgs &a, int mode, cudaStream_t st) {
    const int regime = regime_of(a.num_moduli);
    if (a.row_contig) {
        if (mode == 0) launch_rowcontig<T, 0>(a, regime, st);
        else if (mode == 1) launch_rowcontig<T, 1>(a, regime, st);
        else launch_rowcontig<T, 2>(a, regime, st);
    } else {
        const dim3 sgrid((unsigned)((a.rows + 31) / 32)), sblock(32, 32);
        if (mode == 1) stats_rowstrided_kernel<T, 1><<<sgrid, sblock, 0, st>>>(a);
        if (mode == 2) stats_rowstrided_kernel<T, 2><<<sgrid, sblock, 0, st>>>(a);
        if (mode == 2) launch_rowstrided<T, 2>(a, regime, st);
        else launch_rowstrided<T, 0>(a, regime, st);
    }
    if (mode == 2 && Scalar<T>::cplx) {
        const size_t n16 = a.rows * a.k_pad / 16;
        bound_diff_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>(a.planes[0], a.planes[1], a.planes[2], n16);
    }
}

// mode 0: split with stored shifts, 1: fast (shift + split), 2: accurate stage (i) (s0 + bound planes)
void launch_split(const SplitArgs &a, int dtype, int mode, cudaStream_t st) {
    switch (dtype) {
    case F32: split_typed<float>(a, mode, st); break;
    case F64: split_typed<double>(a, mode, st); break;
    case C32: split_typed<float2>(a, mode, st); break;
    case C64: split_typed<double2>(a, mode, st); break;
    }
}

void launch_finalize_accu_shift(int16_t *sft, const int32_t *cmax, size_t count, int num_moduli, cudaStream_t st) {
    finalize_accu_shift_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(sft, cmax, (int)count, num_moduli);
}

} // namespace g8
