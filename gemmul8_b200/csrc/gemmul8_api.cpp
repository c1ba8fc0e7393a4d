// gemmul8_b200 -- the 16 exported C++ entry points of the reference library (src/gemmul8.cu:95-157), as shims over
// the C ABI.  Host-only code: finds the stream, fills a g8_gemm_desc, forwards.
#include "../../include/gemmul8.hpp"
#include "../../include/gemmul8_c.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#define G8_EXPORT __attribute__((visibility("default")))

namespace {

template <typename T> constexpr int dtype_of() {
    if (std::is_same<T, float>::value) return G8_R32F;
    if (std::is_same<T, double>::value) return G8_R64F;
    if (std::is_same<T, cuFloatComplex>::value) return G8_C32F;
    return G8_C64F;
}

bool want_phase_timing() {
    static const bool on = [] {
        const char *s = std::getenv("GEMMUL8_PHASE_TIMING");
        return s && s[0] == '1';
    }();
    return on;
}

void complain(int code) {
    static std::atomic<int> shown{0};
    if (shown.fetch_add(1) < 8)
        std::fprintf(stderr, "[gemmul8_b200] gemm failed with status %d (%s); C was not updated\n", code,
                     code == G8_STATUS_NOT_SUPPORTED    ? "backend/shape not supported by this build"
                     : code == G8_STATUS_NO_DEVICE_CODE ? "needs an sm_100a (B200) device, there is no fallback path"
                     : code == G8_STATUS_INVALID_VALUE  ? "invalid argument"
                                                        : "CUDA error");
}

template <typename T>
std::vector<double> run(int backend, cudaStream_t stream, cublasOperation_t op_A, cublasOperation_t op_B, size_t m, size_t n, size_t k,
                        const T *alpha, const T *A, size_t lda, const T *B, size_t ldb, const T *beta, T *C, size_t ldc,
                        unsigned num_moduli, bool fastmode, void *work, void *workA, void *workB, bool enA, bool enB, bool skipA,
                        bool skipB) {
    g8_gemm_desc d{};
    d.dtype = dtype_of<T>(), d.backend = backend;
    d.op_A = static_cast<int>(op_A), d.op_B = static_cast<int>(op_B); // CUBLAS_OP_N/T/C == 0/1/2 == G8_OP_*
    d.m = m, d.n = n, d.k = k;
    d.alpha = alpha, d.A = A, d.lda = lda, d.B = B, d.ldb = ldb, d.beta = beta, d.C = C, d.ldc = ldc;
    d.num_moduli = num_moduli, d.fastmode = fastmode;
    d.work = work, d.workA = workA, d.workB = workB;
    d.enable_skip_scalA = enA, d.enable_skip_scalB = enB, d.skip_scalA = skipA, d.skip_scalB = skipB;
    d.stream = stream;
    std::vector<double> t(4, 0.0);
    const int code = g8_gemm(&d, want_phase_timing() ? t.data() : nullptr);
    // the reference silently ignores out-of-range num_moduli (default: break in its switch); keep that quiet too
    if (code != 0 && !(code == G8_STATUS_INVALID_VALUE && (num_moduli < 2 || num_moduli > 20))) complain(code);
    return t;
}

} // namespace

namespace gemmul8 {

#define G8_WORKSIZE(CPLX, BE, BEID)                                                                                                  \
    template <> G8_EXPORT size_t workSize<CPLX, Backend::BE>(size_t m, size_t n, size_t k, unsigned num_moduli, bool enA, bool enB, \
                                                             size_t *wA, size_t *wB) {                                            \
        return g8_work_size(CPLX, BEID, m, n, k, num_moduli, enA, enB, wA, wB);                                                      \
    }
G8_WORKSIZE(false, INT8, G8_BACKEND_INT8)
G8_WORKSIZE(true, INT8, G8_BACKEND_INT8)
G8_WORKSIZE(false, FP8, G8_BACKEND_FP8)
G8_WORKSIZE(true, FP8, G8_BACKEND_FP8)

// gemm<T, INT8>: stream of the cuBLAS handle (src/gemmul8.cu:115-134).  gemm<T, FP8> is intentionally not defined (src/gemmul8.cu:136-139).
#define G8_GEMM(T)                                                                                                                  \
    template <> G8_EXPORT std::vector<double> gemm<T, Backend::INT8>(                                                               \
        cublasHandle_t handle, cublasOperation_t op_A, cublasOperation_t op_B, size_t m, size_t n, size_t k, const T *alpha,        \
        const T *const A, size_t lda, const T *const B, size_t ldb, const T *beta, T *const C, size_t ldc, unsigned num_moduli,     \
        bool fastmode, void *const work, void *const workA, void *const workB, bool enA, bool enB, bool skipA, bool skipB) {        \
        cudaStream_t stream = 0;                                                                                                    \
        cublasGetStream(handle, &stream);                                                                                           \
        return run<T>(G8_BACKEND_INT8, stream, op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, num_moduli, fastmode, work, \
                      workA, workB, enA, enB, skipA, skipB);                                                                        \
    }
G8_GEMM(float)
G8_GEMM(double)
G8_GEMM(cuFloatComplex)
G8_GEMM(cuDoubleComplex)

#define G8_GEMMLT(T, BE, BEID)                                                                                                      \
    template <> G8_EXPORT std::vector<double> gemmLt<T, Backend::BE>(                                                               \
        cublasLtHandle_t, cublasOperation_t op_A, cublasOperation_t op_B, size_t m, size_t n, size_t k, const T *alpha,             \
        const T *const A, size_t lda, const T *const B, size_t ldb, const T *beta, T *const C, size_t ldc, unsigned num_moduli,     \
        bool fastmode, void *const work, void *const workA, void *const workB, bool enA, bool enB, bool skipA, bool skipB,          \
        cudaStream_t stream) {                                                                                                      \
        return run<T>(BEID, stream, op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, num_moduli, fastmode, work, workA,    \
                      workB, enA, enB, skipA, skipB);                                                                               \
    }
G8_GEMMLT(float, INT8, G8_BACKEND_INT8)
G8_GEMMLT(double, INT8, G8_BACKEND_INT8)
G8_GEMMLT(cuFloatComplex, INT8, G8_BACKEND_INT8)
G8_GEMMLT(cuDoubleComplex, INT8, G8_BACKEND_INT8)
G8_GEMMLT(float, FP8, G8_BACKEND_FP8)
G8_GEMMLT(double, FP8, G8_BACKEND_FP8)
G8_GEMMLT(cuFloatComplex, FP8, G8_BACKEND_FP8)
G8_GEMMLT(cuDoubleComplex, FP8, G8_BACKEND_FP8)

} // namespace gemmul8
