// TEST INFRASTRUCTURE ONLY (oracle/): a thin extern "C" shim over the UNMODIFIED
// reference library, so tests/ and `bench.py --impl reference` can drive the
// reference's own public API (gemmul8::gemm / gemmul8::gemmLt / gemmul8::workSize,
// /root/reference/GEMMul8/include/gemmul8.hpp:25-94) from Python through ctypes.
//
// It is compiled together with /root/reference/GEMMul8/src/gemmul8.cu (where it
// lies; sources are never copied into this repo) into oracle/_ref/libgemmul8_ref.so
// by oracle/Makefile.  Nothing in the product path (gemmul8_b200/) may link or
// load this file.
#include <cublasLt.h>
#include <cublas_v2.h>
#include <cuComplex.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <vector>

#include REF_GEMMUL8_HPP
// the reference harness' own matrix generator (testing/make_matrix.hpp:33-82), included where it lies (-I$(REF)/testing)
#include <curand_kernel.h>
#include <type_traits>
#include "make_matrix.hpp"

namespace {
cublasLtHandle_t g_lt = nullptr;
cublasHandle_t g_blas = nullptr;

template <typename T>
int run(int backend, int use_lt, cublasOperation_t opA, cublasOperation_t opB, size_t m, size_t n, size_t k,
        const void *alpha, const void *A, size_t lda, const void *B, size_t ldb, const void *beta, void *C,
        size_t ldc, unsigned num_moduli, bool fastmode, void *work, void *workA, void *workB, bool enA, bool enB,
        bool skipA, bool skipB, cudaStream_t stream, double *timing) {
    std::vector<double> t;
    const T *a = static_cast<const T *>(alpha), *b = static_cast<const T *>(beta);
    const T *pA = static_cast<const T *>(A), *pB = static_cast<const T *>(B);
    T *pC = static_cast<T *>(C);
    if (use_lt) {
        if (!g_lt && cublasLtCreate(&g_lt) != CUBLAS_STATUS_SUCCESS) return 2;
        if (backend == 0)
            t = gemmul8::gemmLt<T, gemmul8::Backend::INT8>(g_lt, opA, opB, m, n, k, a, pA, lda, pB, ldb, b, pC, ldc, num_moduli,
                                                           fastmode, work, workA, workB, enA, enB, skipA, skipB, stream);
        else
            t = gemmul8::gemmLt<T, gemmul8::Backend::FP8>(g_lt, opA, opB, m, n, k, a, pA, lda, pB, ldb, b, pC, ldc, num_moduli,
                                                          fastmode, work, workA, workB, enA, enB, skipA, skipB, stream);
    } else {
        if (backend != 0) return 3; // gemm<T,FP8> is not defined by the reference (gemmul8.cu:136-139)
        if (!g_blas && cublasCreate(&g_blas) != CUBLAS_STATUS_SUCCESS) return 2;
        cublasSetStream(g_blas, stream);
        t = gemmul8::gemm<T, gemmul8::Backend::INT8>(g_blas, opA, opB, m, n, k, a, pA, lda, pB, ldb, b, pC, ldc, num_moduli,
                                                     fastmode, work, workA, workB, enA, enB, skipA, skipB);
    }
    if (timing)
        for (size_t i = 0; i < 4 && i < t.size(); ++i) timing[i] = t[i];
    return 0;
}
} // namespace

extern "C" {

size_t ref_work_size(int is_complex, int backend, size_t m, size_t n, size_t k, unsigned num_moduli, int enA, int enB,
                     size_t *wA, size_t *wB) {
    if (is_complex) {
        if (backend == 0) return gemmul8::workSize<true, gemmul8::Backend::INT8>(m, n, k, num_moduli, enA, enB, wA, wB);
        return gemmul8::workSize<true, gemmul8::Backend::FP8>(m, n, k, num_moduli, enA, enB, wA, wB);
    }
    if (backend == 0) return gemmul8::workSize<false, gemmul8::Backend::INT8>(m, n, k, num_moduli, enA, enB, wA, wB);
    return gemmul8::workSize<false, gemmul8::Backend::FP8>(m, n, k, num_moduli, enA, enB, wA, wB);
}

// dtype: 0 float, 1 double, 2 cuFloatComplex, 3 cuDoubleComplex.  op: 0 N, 1 T, 2 C (== cublasOperation_t).
int ref_gemm(int dtype, int backend, int use_lt, int opA, int opB, size_t m, size_t n, size_t k, const void *alpha,
             const void *A, size_t lda, const void *B, size_t ldb, const void *beta, void *C, size_t ldc,
             unsigned num_moduli, int fastmode, void *work, void *workA, void *workB, int enA, int enB, int skipA,
             int skipB, void *stream, double *timing) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cublasOperation_t oa = static_cast<cublasOperation_t>(opA), ob = static_cast<cublasOperation_t>(opB);
    switch (dtype) {
    case 0: return run<float>(backend, use_lt, oa, ob, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, num_moduli, fastmode, work, workA, workB, enA, enB, skipA, skipB, s, timing);
    case 1: return run<double>(backend, use_lt, oa, ob, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, num_moduli, fastmode, work, workA, workB, enA, enB, skipA, skipB, s, timing);
    case 2: return run<cuFloatComplex>(backend, use_lt, oa, ob, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, num_moduli, fastmode, work, workA, workB, enA, enB, skipA, skipB, s, timing);
    case 3: return run<cuDoubleComplex>(backend, use_lt, oa, ob, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, num_moduli, fastmode, work, workA, workB, enA, enB, skipA, skipB, s, timing);
    }
    return 1;
}

// Synthetic inputs with the reference harness' generator (testing/make_matrix.hpp:33-82), so that `bench.py --impl reference`
// needs nothing from the product library.  Column-major rows x cols, ld = rows; synchronises like the harness does.
int ref_randmat(int dtype, void *X, size_t rows, size_t cols, double phi, unsigned long long seed) {
    switch (dtype) {
    case 0: makemat::randmat<float>(rows, cols, static_cast<float *>(X), phi, seed); break;
    case 1: makemat::randmat<double>(rows, cols, static_cast<double *>(X), phi, seed); break;
    case 2: makemat::randmat<cuFloatComplex>(rows, cols, static_cast<cuFloatComplex *>(X), phi, seed); break;
    case 3: makemat::randmat<cuDoubleComplex>(rows, cols, static_cast<cuDoubleComplex *>(X), phi, seed); break;
    default: return 1;
    }
    return (int)cudaGetLastError();
}

// Plain cuBLASLt TN GEMM exactly as the reference's inner call sets it up (matmult.hpp:41-101,165-169): the "practical tensor
// ceiling" rows of BASELINE.md section 2a.  backend 0: s8 x s8 -> s32 (COMPUTE_32I), 1: e4m3 x e4m3 -> f32 (COMPUTE_32F).
// C[m x n] (ld m) = A^T[k x m] * B[k x n]; heuristic algo #0 with the given workspace, descriptors cached per (backend, shape).
int ref_lt_lowgemm(int backend, size_t m, size_t n, size_t k, const void *A, const void *B, void *C, void *ws, size_t ws_bytes, void *stream) {
    if (!g_lt && cublasLtCreate(&g_lt) != CUBLAS_STATUS_SUCCESS) return 2;
    static cublasLtMatmulDesc_t op = nullptr;
    static cublasLtMatrixLayout_t Ad = nullptr, Bd = nullptr, Cd = nullptr;
    static cublasLtMatmulHeuristicResult_t heur{};
    static size_t cm = 0, cn = 0, ck = 0;
    static int cb = -1;
    if (cm != m || cn != n || ck != k || cb != backend) {
        if (op) { cublasLtMatrixLayoutDestroy(Ad); cublasLtMatrixLayoutDestroy(Bd); cublasLtMatrixLayoutDestroy(Cd); cublasLtMatmulDescDestroy(op); op = nullptr; }
        const auto lowT = backend == 0 ? CUDA_R_8I : CUDA_R_8F_E4M3;
        const auto hiT  = backend == 0 ? CUDA_R_32I : CUDA_R_32F;
        cublasOperation_t ta = CUBLAS_OP_T, tb = CUBLAS_OP_N;
        cublasLtMatmulDescCreate(&op, backend == 0 ? CUBLAS_COMPUTE_32I : CUBLAS_COMPUTE_32F, hiT);
        cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta));
        cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb));
        cublasLtMatrixLayoutCreate(&Ad, lowT, k, m, (int64_t)k);
        cublasLtMatrixLayoutCreate(&Bd, lowT, k, n, (int64_t)k);
        cublasLtMatrixLayoutCreate(&Cd, hiT, m, n, (int64_t)m);
        cublasLtMatmulPreference_t pref;
        cublasLtMatmulPreferenceCreate(&pref);
        cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof(ws_bytes));
        int got = 0;
        cublasLtMatmulAlgoGetHeuristic(g_lt, op, Ad, Bd, Cd, Cd, pref, 1, &heur, &got);
        cublasLtMatmulPreferenceDestroy(pref);
        if (!got) return 4;
        cm = m; cn = n; ck = k; cb = backend;
    }
    const int32_t ione = 1, izero = 0;
    const float fone = 1.0f, fzero = 0.0f;
    const void *one = backend == 0 ? (const void *)&ione : (const void *)&fone, *zero = backend == 0 ? (const void *)&izero : (const void *)&fzero;
    cublasStatus_t st = cublasLtMatmul(g_lt, op, one, A, Ad, B, Bd, zero, C, Cd, C, Cd, &heur.algo, ws, heur.workspaceSize,
                                       static_cast<cudaStream_t>(stream));
    return st == CUBLAS_STATUS_SUCCESS ? 0 : 5;
}

} // extern "C"
