// gemmul8_b200 -- public C++ API, source-compatible with RIKEN-RCCS/GEMMul8 (reference include/gemmul8.hpp:17-94):
// same namespace, enum, template signatures and default arguments, so code written against the reference compiles
// and links against lib/libgemmul8.{a,so} of this repository unchanged.  Every specialisation is a thin shim over the
// C ABI in include/gemmul8_c.h (implemented in gemmul8_b200/csrc/gemmul8_api.cpp).
//
// Differences in behaviour:
//  * the handles are only used to find the stream (gemm: cublasGetStream; gemmLt: the explicit `stream` argument);
//    no cuBLAS / cuBLASLt routine runs on the hot path;
//  * the call is asynchronous; the returned vector {split, gemm, requant, crt} [ns] is all zeros unless the
//    environment variable GEMMUL8_PHASE_TIMING=1 is set (then the call synchronises, as the reference always does);
//  * Backend::FP8 (gemmLt only, like the reference) runs the e4m3 emulation on tcgen05 kind::f8f6f4 for all four types; its
//    accurate-mode shifts may differ by one from the reference's on a floor() boundary (f32 bound product).
#pragma once
#include <cublasLt.h>
#include <cublas_v2.h>
#include <cuComplex.h>
#include <cuda_runtime.h>

#include <cstddef>
#include <vector>

namespace gemmul8 {

enum class Backend { INT8, FP8 };

// Required workspace in bytes (reference include/gemmul8.hpp:25-35).  k <= 2^17; 2 <= num_moduli <= 20.
template <bool is_Complex = false, Backend backend = Backend::INT8>
size_t workSize(size_t m, size_t n, size_t k, unsigned num_moduli, bool enable_skip_scalA = false, bool enable_skip_scalB = false,
                size_t *workSizeA = nullptr, size_t *workSizeB = nullptr);

// C = alpha * op(A) * op(B) + beta * C emulated with INT8 tensor cores (reference include/gemmul8.hpp:41-66).
template <typename T, Backend backend = Backend::INT8>
std::vector<double> gemm(cublasHandle_t handle, cublasOperation_t op_A, cublasOperation_t op_B, size_t m, size_t n, size_t k,
                         const T *alpha, const T *const A, size_t lda, const T *const B, size_t ldb, const T *beta, T *const C,
                         size_t ldc, unsigned num_moduli, bool fastmode, void *const work, void *const workA = nullptr,
                         void *const workB = nullptr, bool enable_skip_scalA = false, bool enable_skip_scalB = false,
                         bool skip_scalA = false, bool skip_scalB = false);

// Same with a cuBLASLt handle and an explicit stream (reference include/gemmul8.hpp:68-94).
template <typename T, Backend backend = Backend::INT8>
std::vector<double> gemmLt(cublasLtHandle_t handle, cublasOperation_t op_A, cublasOperation_t op_B, size_t m, size_t n, size_t k,
                           const T *alpha, const T *const A, size_t lda, const T *const B, size_t ldb, const T *beta, T *const C,
                           size_t ldc, unsigned num_moduli, bool fastmode, void *const work, void *const workA = nullptr,
                           void *const workB = nullptr, bool enable_skip_scalA = false, bool enable_skip_scalB = false,
                           bool skip_scalA = false, bool skip_scalB = false, cudaStream_t stream = 0);

} // namespace gemmul8
