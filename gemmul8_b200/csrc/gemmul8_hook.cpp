// gemmul8_b200 -- LD_PRELOAD shim: cublas{S,D,C,Z}gemm_v2, cublasGemmEx and cublasDestroy_v2 are intercepted and
// routed to the emulator when the GEMMUL8_* environment asks for it; everything else falls through to the real
// cuBLAS found with dlsym(RTLD_NEXT).  Same six symbols and the same 17 environment variables as the reference
// (src/hook.cu:20-38, 846-1055; README.md:302-319); host-only code, nothing here is on the GPU hot path.
//
// Behavioural contract restated from the reference:
//   * GEMMUL8_NUM_MOD_{S,D,C,Z} outside [2, 13|20]  -> native routine               (hook.cu:623-629, 961-1030)
//   * GEMMUL8_FASTMODE_*, GEMMUL8_BACKEND, GEMMUL8_SKIP_SCALE_{A,B} are re-read on every call   (hook.cu:284-310)
//   * GEMMUL8_MAX_{M,N,K,NUM_MOD}, GEMMUL8_MAXWS_BACKEND size the workspaces once per process    (hook.cu:230-281)
//   * one workspace triple (A, B, rest) per cuBLAS handle, grow-only, cudaMallocAsync on the call's stream (hook.cu:331-374)
//   * operand planes are reused when pointer, shape, leading dimension, op, num_moduli, k, type, mode and backend all
//     match the previous call on this handle and the corresponding SKIP switch is on        (hook.cu:81-108, 688-691)
//   * m,n,k <= 0 -> SUCCESS; null A/B/C -> INVALID_VALUE; allocation failure -> ALLOC_FAILED  (hook.cu:616-617, 362-368)
//   * stream switches on a handle are ordered with an event                                 (hook.cu:141-162)
// GEMMUL8_BACKEND=FP8 routes to the FP8 (e4m3) emulation exactly like the reference (hook.cu:567-584); no cuBLASLt handle is
// needed here because the contraction is ours.
#include "../../include/gemmul8.hpp"
#include "../../include/gemmul8_c.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>

#define G8_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

// ------------------------------------------------------------------ environment
unsigned long long env_uint(const char *name, unsigned long long dflt) {
    const char *s = std::getenv(name);
    if (!s || !*s) return dflt;
    char *end = nullptr;
    const unsigned long long v = std::strtoull(s, &end, 10);
    return (end == s) ? dflt : v;
}
bool env_flag(const char *name) {
    const char *s = std::getenv(name);
    return s && std::strcmp(s, "1") == 0;
}
// 0 / "INT8" -> 0, 1 / "FP8" -> 1, (2 / "BOTH" -> 2 when allowed)
int env_backend(const char *name, bool allow_both) {
    const char *s = std::getenv(name);
    if (!s) return 0;
    if (!std::strcmp(s, "1") || !std::strcmp(s, "FP8")) return 1;
    if (allow_both && (!std::strcmp(s, "2") || !std::strcmp(s, "BOTH"))) return 2;
    return 0;
}

struct TypeInfo {
    char tag;          // 'S','D','C','Z'
    int dtype;         // G8_R32F ...
    bool cplx;
    unsigned max_moduli;
    const char *env_num, *env_fast, *native_sym;
};
constexpr TypeInfo kTypes[4] = {
    {'S', G8_R32F, false, 13, "GEMMUL8_NUM_MOD_S", "GEMMUL8_FASTMODE_S", "cublasSgemm_v2"},
    {'D', G8_R64F, false, 20, "GEMMUL8_NUM_MOD_D", "GEMMUL8_FASTMODE_D", "cublasDgemm_v2"},
    {'C', G8_C32F, true, 13, "GEMMUL8_NUM_MOD_C", "GEMMUL8_FASTMODE_C", "cublasCgemm_v2"},
    {'Z', G8_C64F, true, 20, "GEMMUL8_NUM_MOD_Z", "GEMMUL8_FASTMODE_Z", "cublasZgemm_v2"},
};

struct Policy {
    unsigned num_moduli;
    bool fast, keepA, keepB;
    int backend;
    bool emulate;
};
Policy read_policy(const TypeInfo &t) {
    Policy p{};
    p.num_moduli = (unsigned)env_uint(t.env_num, 0);
    p.fast       = env_flag(t.env_fast);
    p.keepA      = env_flag("GEMMUL8_SKIP_SCALE_A");
    p.keepB      = env_flag("GEMMUL8_SKIP_SCALE_B");
    p.backend    = env_backend("GEMMUL8_BACKEND", false);
    p.emulate    = p.num_moduli >= 2 && p.num_moduli <= t.max_moduli && (p.backend == G8_BACKEND_INT8 || p.backend == G8_BACKEND_FP8);
    return p;
}

// process-wide workspace floors used when plane caching is enabled (so cached planes never move)
struct Floors {
    size_t a = 0, b = 0, c = 0;
};
const Floors &floors() {
    static Floors f;
    static std::once_flag once;
    std::call_once(once, [] {
        const size_t mm = env_uint("GEMMUL8_MAX_M", 0), mn = env_uint("GEMMUL8_MAX_N", 0), mk = env_uint("GEMMUL8_MAX_K", 0);
        const unsigned nm = (unsigned)env_uint("GEMMUL8_MAX_NUM_MOD", 2);
        const bool cplx   = env_uint("GEMMUL8_NUM_MOD_Z", 0) > 0 || env_uint("GEMMUL8_NUM_MOD_C", 0) > 0;
        const int which   = env_backend("GEMMUL8_MAXWS_BACKEND", true);
        for (int be = 0; be < 2; ++be) {
            if (!(which == 2 || which == be)) continue;
            size_t wa = 0, wb = 0;
            const size_t tot = g8_work_size(cplx, be, mm, mn, mk, nm, 1, 1, &wa, &wb);
            f.a = std::max(f.a, wa), f.b = std::max(f.b, wb);
            f.c = std::max(f.c, tot > wa + wb ? tot - wa - wb : 0);
        }
    });
    return f;
}

// ------------------------------------------------------------------ per-handle session
struct Pool {
    void *ptr = nullptr;
    size_t bytes = 0;
    // grow-only; returns nullptr on failure
    void *reserve(size_t want, cudaStream_t s, const char *what, cublasStatus_t &st) {
        st = CUBLAS_STATUS_SUCCESS;
        if (want == 0) return nullptr;
        if (ptr && bytes >= want) return ptr;
        if (ptr) {
            if (cudaFreeAsync(ptr, s) != cudaSuccess) {
                std::fprintf(stderr, "[gemmul8 hook] cudaFreeAsync(%s) failed\n", what);
                st = CUBLAS_STATUS_INTERNAL_ERROR;
                return nullptr;
            }
            ptr = nullptr, bytes = 0;
        }
        void *p = nullptr;
        const cudaError_t e = cudaMallocAsync(&p, want, s);
        if (e != cudaSuccess) {
            std::fprintf(stderr, "[gemmul8 hook] cudaMallocAsync(%s, %zu bytes) failed: %s\n", what, want, cudaGetErrorString(e));
            st = CUBLAS_STATUS_ALLOC_FAILED;
            return nullptr;
        }
        ptr = p, bytes = want;
        return ptr;
    }
    void release(cudaStream_t s, bool have_stream) {
        if (!ptr) return;
        if (!have_stream || cudaFreeAsync(ptr, s) != cudaSuccess) {
            if (have_stream) cudaStreamSynchronize(s);
            cudaFree(ptr);
        }
        ptr = nullptr, bytes = 0;
    }
};

struct Fingerprint { // what must be unchanged for cached planes to be valid
    unsigned num_moduli = 0;
    size_t k = 0;
    char tag = 0;
    bool fast = false;
    int backend = -1;
    const void *A = nullptr, *B = nullptr;
    void *poolA = nullptr, *poolB = nullptr;
    size_t m = 0, n = 0, lda = 0, ldb = 0;
    int opA = -1, opB = -1;
    bool keepA = false, keepB = false; // the workspace layout (shift vectors, accurate-mode bound planes) depends on these
};

struct Session {
    std::mutex mtx;
    Pool a, b, c;
    Fingerprint last;
    cudaStream_t stream = nullptr;
    bool stream_known   = false;
    cudaEvent_t fence   = nullptr;
};

std::mutex g_sessions_mtx;
std::unordered_map<cublasHandle_t, std::shared_ptr<Session>> g_sessions;

std::shared_ptr<Session> session_of(cublasHandle_t h) {
    std::lock_guard<std::mutex> g(g_sessions_mtx);
    auto &s = g_sessions[h];
    if (!s) s = std::make_shared<Session>();
    return s;
}

// consecutive calls on one handle but different streams share the workspaces: order them with an event
cublasStatus_t order_streams(Session &s, cudaStream_t now) {
    if (!s.stream_known) {
        s.stream = now, s.stream_known = true;
        return CUBLAS_STATUS_SUCCESS;
    }
    if (s.stream == now) return CUBLAS_STATUS_SUCCESS;
    if (!s.fence && cudaEventCreateWithFlags(&s.fence, cudaEventDisableTiming) != cudaSuccess) return CUBLAS_STATUS_INTERNAL_ERROR;
    if (cudaEventRecord(s.fence, s.stream) != cudaSuccess) return CUBLAS_STATUS_INTERNAL_ERROR;
    if (cudaStreamWaitEvent(now, s.fence, 0) != cudaSuccess) return CUBLAS_STATUS_INTERNAL_ERROR;
    s.stream = now;
    return CUBLAS_STATUS_SUCCESS;
}

template <typename Fn> Fn native(const char *sym) { return reinterpret_cast<Fn>(dlsym(RTLD_NEXT, sym)); }
constexpr cublasStatus_t kFallBack = static_cast<cublasStatus_t>(-1); // internal: "run the native routine instead"

// ------------------------------------------------------------------ the emulated call
cublasStatus_t emulate(const TypeInfo &t, const Policy &p, cublasHandle_t handle, cublasOperation_t opA, cublasOperation_t opB, int m,
                       int n, int k, const void *alpha, const void *A, int lda, const void *B, int ldb, const void *beta, void *C,
                       int ldc) {
    auto sp = session_of(handle);
    std::lock_guard<std::mutex> lock(sp->mtx);
    Session &s = *sp;

    cudaStream_t stream = nullptr;
    cublasStatus_t st   = cublasGetStream(handle, &stream);
    if (st != CUBLAS_STATUS_SUCCESS) return st;
    if ((st = order_streams(s, stream)) != CUBLAS_STATUS_SUCCESS) return st;

    size_t needA = 0, needB = 0;
    const size_t total = g8_work_size(t.cplx, p.backend, m, n, k, p.num_moduli, p.keepA, p.keepB, &needA, &needB);
    if (total < needA + needB) return CUBLAS_STATUS_INVALID_VALUE;
    size_t needC = total - needA - needB;
    if (p.keepA || p.keepB) {
        const Floors &f = floors();
        if (p.keepA) needA = std::max(needA, f.a);
        if (p.keepB) needB = std::max(needB, f.b);
        needC = std::max(needC, f.c);
    }
    void *wA = s.a.reserve(needA, stream, "workA", st);
    if (st != CUBLAS_STATUS_SUCCESS) return st;
    void *wB = s.b.reserve(needB, stream, "workB", st);
    if (st != CUBLAS_STATUS_SUCCESS) return st;
    void *wC = s.c.reserve(needC, stream, "workC", st);
    if (st != CUBLAS_STATUS_SUCCESS) return st;

    Fingerprint now;
    now.num_moduli = p.num_moduli, now.k = (size_t)k, now.tag = t.tag, now.fast = p.fast, now.backend = p.backend;
    now.A = A, now.B = B, now.poolA = wA, now.poolB = wB;
    now.m = (size_t)m, now.n = (size_t)n, now.lda = (size_t)lda, now.ldb = (size_t)ldb, now.opA = (int)opA, now.opB = (int)opB;
    now.keepA = p.keepA, now.keepB = p.keepB;
    const Fingerprint &was = s.last;
    const bool same_core   = was.num_moduli == now.num_moduli && was.k == now.k && was.tag == now.tag && was.fast == now.fast &&
                           was.backend == now.backend && was.keepA == now.keepA && was.keepB == now.keepB;
    const bool reuseA = same_core && p.keepA && was.poolA == now.poolA && was.A == now.A && was.m == now.m && was.lda == now.lda &&
                        was.opA == now.opA;
    const bool reuseB = same_core && p.keepB && was.poolB == now.poolB && was.B == now.B && was.n == now.n && was.ldb == now.ldb &&
                        was.opB == now.opB;

    g8_gemm_desc d{};
    d.dtype = t.dtype, d.backend = p.backend, d.op_A = (int)opA, d.op_B = (int)opB;
    d.m = (size_t)m, d.n = (size_t)n, d.k = (size_t)k;
    d.alpha = alpha, d.A = A, d.lda = (size_t)lda, d.B = B, d.ldb = (size_t)ldb, d.beta = beta, d.C = C, d.ldc = (size_t)ldc;
    d.num_moduli = p.num_moduli, d.fastmode = p.fast;
    d.work = wC, d.workA = wA, d.workB = wB;
    d.enable_skip_scalA = p.keepA, d.enable_skip_scalB = p.keepB, d.skip_scalA = reuseA, d.skip_scalB = reuseB;
    d.stream = stream;
    const int code = g8_gemm(&d, nullptr);
    if (code != 0) {
        s.last = Fingerprint{};
        // argument / capability rejections leave C untouched: let the caller run the native routine instead of failing the application
        if (code == G8_STATUS_INVALID_VALUE || code == G8_STATUS_NOT_SUPPORTED || code == G8_STATUS_NO_DEVICE_CODE) return kFallBack;
        std::fprintf(stderr, "[gemmul8 hook] emulated %cGEMM failed with CUDA status %d\n", t.tag, code);
        return CUBLAS_STATUS_EXECUTION_FAILED;
    }
    s.last = now;
    return CUBLAS_STATUS_SUCCESS;
}

// Emulation is attempted only where it is exact and can run: an sm_100a device and k within the accumulation bound of the backend
// (2^17 for INT8, 2^16 for FP8: include/gemmul8_c.h).  Everything else goes to the native routine, as if the hook were not there.
bool can_emulate(const Policy &p, int k) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || !g8_device_supported(dev)) return false;
    return (size_t)k <= (p.backend == G8_BACKEND_FP8 ? (size_t(1) << 16) : (size_t(1) << 17));
}

template <typename T>
cublasStatus_t gemm_v2(const TypeInfo &t, cublasHandle_t handle, cublasOperation_t opA, cublasOperation_t opB, int m, int n, int k,
                       const T *alpha, const T *A, int lda, const T *B, int ldb, const T *beta, T *C, int ldc) {
    if (m <= 0 || n <= 0 || k <= 0) return CUBLAS_STATUS_SUCCESS;
    if (!A || !B || !C) return CUBLAS_STATUS_INVALID_VALUE;
    const Policy p = read_policy(t);
    if (p.emulate && can_emulate(p, k)) {
        const cublasStatus_t st = emulate(t, p, handle, opA, opB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
        if (st != kFallBack) return st;
    }
    using Fn = cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const T *, const T *, int,
                                  const T *, int, const T *, T *, int);
    Fn fn = native<Fn>(t.native_sym);
    return fn ? fn(handle, opA, opB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc) : CUBLAS_STATUS_NOT_INITIALIZED;
}

} // namespace

G8_EXPORT cublasStatus_t cublasSgemm_v2(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                                        const float *alpha, const float *A, int lda, const float *B, int ldb, const float *beta,
                                        float *C, int ldc) {
    return gemm_v2<float>(kTypes[0], h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
G8_EXPORT cublasStatus_t cublasDgemm_v2(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                                        const double *alpha, const double *A, int lda, const double *B, int ldb, const double *beta,
                                        double *C, int ldc) {
    return gemm_v2<double>(kTypes[1], h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
G8_EXPORT cublasStatus_t cublasCgemm_v2(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                                        const cuComplex *alpha, const cuComplex *A, int lda, const cuComplex *B, int ldb,
                                        const cuComplex *beta, cuComplex *C, int ldc) {
    return gemm_v2<cuComplex>(kTypes[2], h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
G8_EXPORT cublasStatus_t cublasZgemm_v2(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                                        const cuDoubleComplex *alpha, const cuDoubleComplex *A, int lda, const cuDoubleComplex *B,
                                        int ldb, const cuDoubleComplex *beta, cuDoubleComplex *C, int ldc) {
    return gemm_v2<cuDoubleComplex>(kTypes[3], h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

// cublasGemmEx: only the four "plain" S/D/C/Z type combinations are candidates (hook.cu:961-1030)
G8_EXPORT cublasStatus_t cublasGemmEx(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                                      const void *alpha, const void *A, cudaDataType At, int lda, const void *B, cudaDataType Bt, int ldb,
                                      const void *beta, void *C, cudaDataType Ct, int ldc, cublasComputeType_t ct, cublasGemmAlgo_t algo) {
    if (m <= 0 || n <= 0 || k <= 0) return CUBLAS_STATUS_SUCCESS;
    if (!A || !B || !C) return CUBLAS_STATUS_INVALID_VALUE;
    const TypeInfo *t = nullptr;
    if (At == Bt && Bt == Ct) {
        if (ct == CUBLAS_COMPUTE_32F && At == CUDA_R_32F) t = &kTypes[0];
        else if (ct == CUBLAS_COMPUTE_64F && At == CUDA_R_64F) t = &kTypes[1];
        else if (ct == CUBLAS_COMPUTE_32F && At == CUDA_C_32F) t = &kTypes[2];
        else if (ct == CUBLAS_COMPUTE_64F && At == CUDA_C_64F) t = &kTypes[3];
    }
    if (t) {
        const Policy p = read_policy(*t);
        if (p.emulate && can_emulate(p, k)) {
            const cublasStatus_t st = emulate(*t, p, h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
            if (st != kFallBack) return st;
        }
    }
    using Fn = cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const void *, const void *,
                                  cudaDataType, int, const void *, cudaDataType, int, const void *, void *, cudaDataType, int,
                                  cublasComputeType_t, cublasGemmAlgo_t);
    static Fn fn = native<Fn>("cublasGemmEx");
    return fn ? fn(h, ta, tb, m, n, k, alpha, A, At, lda, B, Bt, ldb, beta, C, Ct, ldc, ct, algo) : CUBLAS_STATUS_NOT_INITIALIZED;
}

// release the handle's workspaces before the real destroy (hook.cu:376-462, 846-856)
G8_EXPORT cublasStatus_t cublasDestroy_v2(cublasHandle_t h) {
    std::shared_ptr<Session> sp;
    {
        std::lock_guard<std::mutex> g(g_sessions_mtx);
        auto it = g_sessions.find(h);
        if (it != g_sessions.end()) {
            sp = it->second;
            g_sessions.erase(it);
        }
    }
    if (sp) {
        std::lock_guard<std::mutex> lock(sp->mtx);
        cudaStream_t s = sp->stream;
        bool have      = sp->stream_known;
        if (!have && cublasGetStream(h, &s) == CUBLAS_STATUS_SUCCESS) have = true;
        if (!have) cudaDeviceSynchronize();
        sp->a.release(s, have), sp->b.release(s, have), sp->c.release(s, have);
        if (sp->fence) cudaEventDestroy(sp->fence);
    }
    using Fn = cublasStatus_t (*)(cublasHandle_t);
    static Fn fn = native<Fn>("cublasDestroy_v2");
    return fn ? fn(h) : CUBLAS_STATUS_NOT_INITIALIZED;
}
