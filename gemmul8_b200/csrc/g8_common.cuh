// gemmul8_b200 -- shared device/host definitions for the Ozaki-II hot path (sm_100a only).
//
// Layout contract (kept identical to the reference so that caller-owned workspaces, the
// skip-scaling cache and plane-level parity tests line up; reference gemmul8_real.hpp:95-107,
// gemmul8_complex.hpp:95-118):
//   A_lo : planes of int8, each k_pad x m_pad, K-major (row r of op(A) is k_pad contiguous bytes)
//   B_lo : planes of int8, each k_pad x n,     K-major (column c of op(B) is k_pad contiguous bytes)
//   sftA : int16[m_pad], sftB : int16[pad256(n)]  (shift exponents, stored NEGATED)
//   C_mid: num_moduli planes, m_pad x n column-major; int8 (real) or int8x2 {re,im} (complex)
// with k_pad = pad256(k), m_pad = pad256(m).
#pragma once
#include <cuda_runtime.h>
#include <cuComplex.h>
#include <cstddef>
#include <cstdint>

#define G8_MAX_MODULI 20
#define G8_MAX_PEERS 8 // GPUs of one NVSwitch box

// ---- constant tables: one host copy (g8h_*) and one __constant__ copy (g8d_*) per translation unit
#define G8_TQ static const
#define G8_TN(name) g8h_##name
#include "g8_tables.h"
#undef G8_TQ
#undef G8_TN
#define G8_TQ static __device__ __constant__
#define G8_TN(name) g8d_##name
#include "g8_tables.h"
#undef G8_TQ
#undef G8_TN

namespace g8 {

enum Backend : int { INT8 = 0, FP8 = 1 };
enum DType : int { F32 = 0, F64 = 1, C32 = 2, C64 = 3 };
enum Op : int { OP_N = 0, OP_T = 1, OP_C = 2 }; // == cublasOperation_t values

// common.hpp:15-27 of the reference: representation thresholds of the scaled operand
struct Thresholds {
    int P_is_double, S, M;
};
__host__ __device__ inline Thresholds thresholds(int backend) {
    return backend == INT8 ? Thresholds{6, 7, 15} : Thresholds{5, 5, 12};
}

__host__ __device__ inline size_t pad256(size_t x) { return 256 * ((x + 255) / 256); }

inline void *align256(void *p) {
    uintptr_t x = reinterpret_cast<uintptr_t>(p);
    return reinterpret_cast<void *>((x + 255) & ~uintptr_t(255));
}

// ---------------------------------------------------------------------------------------------
// symmetric residue helpers (device).  Any correct method yields the unique representative in
// [-(p-1)/2, (p-1)/2] for odd p, so these are bit-compatible with reference mod.hpp:8-37 without
// sharing its code.  p == 256 keeps +128, which the int8 store turns into -128 exactly as the
// reference's static_cast<int8_t> does.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t sym_wrap(int32_t r, int32_t p) {
    const int32_t h = p >> 1;
    r               = (r > h) ? r - p : r;
    r               = (r < -h) ? r + p : r;
    return r;
}
// |x| < 2^31, pinv32 = floor(2^32/p): q = mulhi(x, pinv32) under-estimates x/p by < 1 + |x|/2^32,
// so r = x - p*q lies in (-p/2, 3p/2) and one wrap lands in the symmetric range.
__device__ __forceinline__ int32_t mod_i32(int32_t x, int32_t p, int32_t pinv32) {
    return sym_wrap(x - p * __mulhi(x, pinv32), p);
}
// |x| < 2^63, pinv64 = floor(2^64/p): same bound with 2^64
__device__ __forceinline__ int32_t mod_i64(int64_t x, int32_t p, int64_t pinv64) {
    return sym_wrap(int32_t(x - int64_t(p) * __mul64hi(x, pinv64)), p);
}

} // namespace g8
