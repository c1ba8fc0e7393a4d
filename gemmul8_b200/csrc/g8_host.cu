// gemmul8_b200 -- emulated GEMM on HOST buffers (include/gemmul8_c.h: g8_host_plan_*, g8_gemm_host).
//
// The reference's API takes device pointers only (include/gemmul8.hpp:41-94); an application whose matrices live in host memory wraps
// the call in three bulk copies and PCIe dominates (1.5 GB at ~55 GB/s = 28 ms against ~7 ms of GPU work for DGEMM 8192^3).  The
// emulation is separable along the columns of op(B) / C -- the shift of column c needs only that column (plus all of A) and C[:, c]
// needs only column c of the residue planes -- so B and C stream in column chunks and the copies overlap the stage kernels on three
// streams:
//     h2d  : A .......| B[:,0] | B[:,1] | B[:,2] | ...
//     comp :          | splitA | chunk 0: splitB, GEMMs (all moduli), CRT | chunk 1 ... |
//     d2h  :                                                          | C[:,0] | C[:,1] | ...
// Accurate mode needs the row maxima of the bound product over ALL columns before A can be split: the B side (bound planes, bound
// GEMM, final B shifts, B split) runs chunk-wise while B streams in, then A is split and GEMMs / CRT / D2H are pipelined.
// Every number is produced by the same stage kernels as g8_gemm (contract(), bound_gemm(), launch_split, launch_crt), so the result
// is bit-identical to the monolithic call.  All four types, both backends, all op combinations.
#include "g8_internal.cuh"
#include "../../include/gemmul8_c.h"

#include <algorithm>
#include <new>

namespace g8 {

struct HostPlan {
    int dtype, backend, opA, opB, fast;
    size_t m, n, k, chunk;
    unsigned N;
    bool cplx;
    size_t esz;                       // bytes per element
    size_t k_pad, m_pad, n_pad, sizeA, sizeB, sizeC, mid;
    unsigned nmat, sets;
    int8_t *A_lo = nullptr, *B_lo = nullptr, *C_mid = nullptr, *scratch = nullptr;
    size_t scratch_bytes = 0;
    int16_t *sftA = nullptr, *sftB = nullptr;
    int32_t *maxes = nullptr;         // rowmax[m_pad] | colmax[n_pad]
    void *dA = nullptr, *dB[2] = {nullptr, nullptr}, *dC[2] = {nullptr, nullptr};
    size_t dA_elems = 0;
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    cudaEvent_t evA = nullptr, evB[2] = {nullptr, nullptr}, evBfree[2] = {nullptr, nullptr}, evC[2] = {nullptr, nullptr},
                evCin[2] = {nullptr, nullptr}, evCfree[2] = {nullptr, nullptr}, evStart = nullptr, evDone = nullptr;
    int device = 0;
};

#define G8_TRY(x)                              \
    do {                                       \
        cudaError_t _e = (x);                  \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

static void destroy(HostPlan *p) {
    if (!p) return;
    for (void *q : {(void *)p->A_lo, (void *)p->B_lo, (void *)p->C_mid, (void *)p->scratch, (void *)p->sftA, (void *)p->sftB, (void *)p->maxes, p->dA,
                    p->dB[0], p->dB[1], p->dC[0], p->dC[1]})
        if (q) cudaFree(q);
    for (cudaStream_t s : {p->s_h2d, p->s_comp, p->s_d2h})
        if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : {p->evA, p->evB[0], p->evB[1], p->evBfree[0], p->evBfree[1], p->evC[0], p->evC[1], p->evCin[0], p->evCin[1], p->evCfree[0],
                          p->evCfree[1], p->evStart, p->evDone})
        if (e) cudaEventDestroy(e);
    delete p;
}

static int create(HostPlan **out, int dtype, int backend, int opA, int opB, size_t m, size_t n, size_t k, unsigned N, int fast, size_t chunk) {
    if (!out || dtype < F32 || dtype > C64 || (backend != INT8 && backend != FP8) || opA < 0 || opA > 2 || opB < 0 || opB > 2) return G8_STATUS_INVALID_VALUE;
    if (N < 2 || N > G8_MAX_MODULI || m == 0 || n == 0 || k == 0) return G8_STATUS_INVALID_VALUE;
    if (k > (size_t(1) << (backend == FP8 ? 16 : 17))) return G8_STATUS_INVALID_VALUE;
    if (!device_supported_cached()) return G8_STATUS_NO_DEVICE_CODE;
    HostPlan *p = new (std::nothrow) HostPlan();
    if (!p) return (int)cudaErrorMemoryAllocation;
    p->dtype = dtype, p->backend = backend, p->opA = opA, p->opB = opB, p->fast = fast, p->m = m, p->n = n, p->k = k, p->N = N;
    p->cplx = dtype >= C32;
    p->esz  = (dtype == F32 ? 4 : dtype == F64 ? 8 : dtype == C32 ? 8 : 16);
    p->chunk = std::max<size_t>(256, std::min(chunk ? (chunk + 255) / 256 * 256 : size_t(1024), pad256(n))); // whole 256-column tiles
    p->k_pad = pad256(k), p->m_pad = pad256(m), p->n_pad = pad256(n);
    p->sizeA = p->k_pad * p->m_pad, p->sizeB = p->k_pad * n, p->sizeC = p->m_pad * n;
    p->nmat = num_planes(backend, N), p->sets = p->cplx ? 3 : 1;
    p->mid  = (backend == INT8 ? 1 : 2) * (p->cplx ? 2 : 1);
    cudaGetDevice(&p->device);
    auto fail = [&](int code) {
        destroy(p);
        return code;
    };
#define G8_ALLOC(ptr, bytes)                                                               \
    if (cudaMalloc(reinterpret_cast<void **>(&(ptr)), (bytes)) != cudaSuccess) return fail((int)cudaErrorMemoryAllocation)
    G8_ALLOC(p->A_lo, p->sizeA * p->nmat * p->sets);
    G8_ALLOC(p->B_lo, p->sizeB * p->nmat * p->sets);
    G8_ALLOC(p->C_mid, p->mid * p->sizeC * N);
    G8_ALLOC(p->sftA, sizeof(int16_t) * p->m_pad);
    G8_ALLOC(p->sftB, sizeof(int16_t) * p->n_pad);
    G8_ALLOC(p->maxes, sizeof(int32_t) * (p->m_pad + p->n_pad));
    // per-product residues of the multi-product paths, for ALL moduli of one chunk (one contraction launch per chunk)
    const size_t prods = backend == FP8 ? (p->cplx ? 9 : 3) * sizeof(int16_t) : (p->cplx ? 3 : 0);
    p->scratch_bytes   = prods * p->m_pad * p->chunk * N;
    if (p->scratch_bytes) G8_ALLOC(p->scratch, p->scratch_bytes);
    const size_t rowsA = opA == OP_N ? m : k, colsA = opA == OP_N ? k : m;
    p->dA_elems        = rowsA * colsA;
    G8_ALLOC(p->dA, p->dA_elems * p->esz);
    for (int i = 0; i < 2; ++i) {
        G8_ALLOC(p->dB[i], k * p->chunk * p->esz); // chunk columns of op(B), stored compactly (ld = k for op N, ld = chunk columns for op T/C)
        G8_ALLOC(p->dC[i], m * p->chunk * p->esz);
    }
#undef G8_ALLOC
    cudaMemset(p->sftA, 0, sizeof(int16_t) * p->m_pad), cudaMemset(p->sftB, 0, sizeof(int16_t) * p->n_pad);
    for (cudaStream_t *s : {&p->s_h2d, &p->s_comp, &p->s_d2h})
        if (cudaStreamCreateWithFlags(s, cudaStreamNonBlocking) != cudaSuccess) return fail((int)cudaErrorUnknown);
    for (cudaEvent_t *e : {&p->evA, &p->evB[0], &p->evB[1], &p->evBfree[0], &p->evBfree[1], &p->evC[0], &p->evC[1], &p->evCin[0], &p->evCin[1],
                           &p->evCfree[0], &p->evCfree[1], &p->evStart, &p->evDone})
        if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return fail((int)cudaErrorUnknown);
    *out = p;
    return 0;
}

static bool scalar_is_zero(const void *s, int dtype) {
    switch (dtype) {
    case F32: return *static_cast<const float *>(s) == 0.0f;
    case F64: return *static_cast<const double *>(s) == 0.0;
    case C32: return static_cast<const float *>(s)[0] == 0.0f && static_cast<const float *>(s)[1] == 0.0f;
    default: return static_cast<const double *>(s)[0] == 0.0 && static_cast<const double *>(s)[1] == 0.0;
    }
}

static int run(HostPlan &p, const void *alpha, const void *hA, size_t lda, const void *hB, size_t ldb, const void *beta, void *hC, size_t ldc,
               cudaStream_t user) {
    const size_t m = p.m, n = p.n, k = p.k, esz = p.esz;
    const unsigned N = p.N;
    const size_t rowsA = p.opA == OP_N ? m : k, colsA = p.opA == OP_N ? k : m;
    if (!alpha || !beta || !hA || !hB || !hC || lda < rowsA || ldb < (p.opB == OP_N ? k : n) || ldc < m) return G8_STATUS_INVALID_VALUE;
    const bool need_c_in = !scalar_is_zero(beta, p.dtype);
    int32_t *rowmax = p.maxes, *colmax = p.maxes + p.m_pad;

    // order the three streams after the caller's stream
    G8_TRY(cudaEventRecord(p.evStart, user));
    for (cudaStream_t s : {p.s_h2d, p.s_comp, p.s_d2h}) G8_TRY(cudaStreamWaitEvent(s, p.evStart, 0));

    // ---- A: copy, then (fast) shift + split or (accurate) s0 + bound planes ----
    G8_TRY(cudaMemcpy2DAsync(p.dA, rowsA * esz, hA, lda * esz, rowsA * esz, colsA, cudaMemcpyHostToDevice, p.s_h2d));
    G8_TRY(cudaEventRecord(p.evA, p.s_h2d));
    G8_TRY(cudaStreamWaitEvent(p.s_comp, p.evA, 0));
    const SplitArgs sa = make_split_args(1, p.opA, m, k, p.dA, rowsA, N, p.sftA, p.A_lo, p.sizeA, p.nmat, p.backend);
    SplitArgs ea = sa; // accurate stage (i): bound planes alias plane 0 of every plane set
    for (int g = 0; g < 3; ++g) ea.planes[g] = p.A_lo + g * p.sizeA;
    if (p.fast) {
        launch_split(sa, p.dtype, 1, p.s_comp);
    } else {
        G8_TRY(cudaMemsetAsync(p.maxes, 0, sizeof(int32_t) * (p.m_pad + p.n_pad), p.s_comp));
        launch_split(ea, p.dtype, 2, p.s_comp);
    }

    const size_t W = p.chunk;
    const int nchunks = (int)((n + W - 1) / W);
    bool b_used[2] = {false, false}, c_used[2] = {false, false};

    auto b_side = [&](int ci) -> int {
        const int i = ci & 1;
        const size_t c0 = (size_t)ci * W, nc = std::min(W, n - c0);
        if (b_used[i]) G8_TRY(cudaStreamWaitEvent(p.s_h2d, p.evBfree[i], 0));
        size_t ldd; // leading dimension of the device chunk
        if (p.opB == OP_N) { // columns c0.. of the stored k x n matrix: contiguous column block
            ldd = k;
            G8_TRY(cudaMemcpy2DAsync(p.dB[i], k * esz, static_cast<const char *>(hB) + c0 * ldb * esz, ldb * esz, k * esz, nc, cudaMemcpyHostToDevice, p.s_h2d));
        } else { // rows c0.. of the stored n x k matrix: a row block, gathered column by column
            ldd = nc;
            G8_TRY(cudaMemcpy2DAsync(p.dB[i], nc * esz, static_cast<const char *>(hB) + c0 * esz, ldb * esz, nc * esz, k, cudaMemcpyHostToDevice, p.s_h2d));
        }
        G8_TRY(cudaEventRecord(p.evB[i], p.s_h2d));
        G8_TRY(cudaStreamWaitEvent(p.s_comp, p.evB[i], 0));
        int8_t *planes = p.B_lo + c0 * p.k_pad;
        const SplitArgs sb = make_split_args(0, p.opB, nc, k, p.dB[i], ldd, N, p.sftB + c0, planes, p.sizeB, p.nmat, p.backend);
        if (p.fast) {
            launch_split(sb, p.dtype, 1, p.s_comp);
        } else {
            SplitArgs eb = sb;
            for (int g = 0; g < 3; ++g) eb.planes[g] = planes + g * p.sizeB;
            launch_split(eb, p.dtype, 2, p.s_comp);
            if (int e = bound_gemm(p.cplx, p.backend, p.A_lo, p.sizeA, planes, p.sizeB, m, nc, p.m_pad, p.k_pad, k, rowmax, colmax + c0, p.s_comp)) return e;
            launch_finalize_accu_shift(p.sftB + c0, colmax + c0, nc, (int)N, p.s_comp, p.backend);
            launch_split(sb, p.dtype, 0, p.s_comp);
        }
        G8_TRY(cudaEventRecord(p.evBfree[i], p.s_comp));
        b_used[i] = true;
        return 0;
    };

    auto c_side = [&](int ci) -> int {
        const int i = ci & 1;
        const size_t c0 = (size_t)ci * W, nc = std::min(W, n - c0);
        ContractArgs ca{};
        ca.cplx = p.cplx, ca.backend = p.backend, ca.N = N, ca.m = m, ca.ncols = nc, ca.m_pad = p.m_pad, ca.k_pad = p.k_pad;
        ca.A_lo = p.A_lo, ca.B_lo = p.B_lo + c0 * p.k_pad, ca.sizeA = p.sizeA, ca.sizeB = p.sizeB, ca.set_planes = p.nmat;
        ca.C_mid = p.C_mid + c0 * p.m_pad * p.mid, ca.mid_plane_stride = p.sizeC, ca.scratch = p.scratch, ca.scratch_avail = p.scratch_bytes;
        if (int e = contract(ca, p.s_comp)) return e;
        if (c_used[i]) G8_TRY(cudaStreamWaitEvent(p.s_comp, p.evCfree[i], 0)); // the previous D2H of this buffer has finished
        if (need_c_in) {
            if (c_used[i]) G8_TRY(cudaStreamWaitEvent(p.s_h2d, p.evCfree[i], 0));
            G8_TRY(cudaMemcpy2DAsync(p.dC[i], m * esz, static_cast<const char *>(hC) + c0 * ldc * esz, ldc * esz, m * esz, nc, cudaMemcpyHostToDevice, p.s_h2d));
            G8_TRY(cudaEventRecord(p.evCin[i], p.s_h2d));
            G8_TRY(cudaStreamWaitEvent(p.s_comp, p.evCin[i], 0));
        }
        CrtArgs c{};
        c.C_mid = ca.C_mid, c.ldmid = p.m_pad, c.plane_stride = p.sizeC, c.m = m, c.n = nc, c.num_moduli = (int)N;
        c.C = p.dC[i], c.ldc = m, c.sftA = p.sftA, c.sftB = p.sftB + c0, c.alpha = alpha, c.beta = beta, c.backend = p.backend;
        if (int e = launch_crt(c, p.dtype, p.s_comp)) return e;
        G8_TRY(cudaEventRecord(p.evC[i], p.s_comp));
        G8_TRY(cudaStreamWaitEvent(p.s_d2h, p.evC[i], 0));
        G8_TRY(cudaMemcpy2DAsync(static_cast<char *>(hC) + c0 * ldc * esz, ldc * esz, p.dC[i], m * esz, m * esz, nc, cudaMemcpyDeviceToHost, p.s_d2h));
        G8_TRY(cudaEventRecord(p.evCfree[i], p.s_d2h));
        c_used[i] = true;
        return 0;
    };

    if (p.fast) {
        for (int ci = 0; ci < nchunks; ++ci) {
            if (int e = b_side(ci)) return e;
            if (int e = c_side(ci)) return e;
        }
    } else {
        for (int ci = 0; ci < nchunks; ++ci)
            if (int e = b_side(ci)) return e;
        launch_finalize_accu_shift(p.sftA, rowmax, m, (int)N, p.s_comp, p.backend);
        launch_split(sa, p.dtype, 0, p.s_comp);
        for (int ci = 0; ci < nchunks; ++ci)
            if (int e = c_side(ci)) return e;
    }
    // the caller's stream continues after the last D2H (and after the helper streams are idle)
    G8_TRY(cudaEventRecord(p.evDone, p.s_d2h));
    G8_TRY(cudaStreamWaitEvent(user, p.evDone, 0));
    G8_TRY(cudaEventRecord(p.evDone, p.s_comp));
    G8_TRY(cudaStreamWaitEvent(user, p.evDone, 0));
    G8_TRY(cudaEventRecord(p.evDone, p.s_h2d));
    G8_TRY(cudaStreamWaitEvent(user, p.evDone, 0));
    return (int)cudaPeekAtLastError();
}

} // namespace g8

using namespace g8;

extern "C" {

__attribute__((visibility("default"))) int g8_host_plan_create(g8_host_plan **plan, int dtype, int backend, int op_A, int op_B, size_t m, size_t n, size_t k,
                                                                unsigned num_moduli, int fastmode, size_t chunk_cols) {
    HostPlan *p = nullptr;
    const int e = create(&p, dtype, backend, op_A, op_B, m, n, k, num_moduli, fastmode != 0, chunk_cols);
    if (e == 0) *plan = reinterpret_cast<g8_host_plan *>(p);
    return e;
}

__attribute__((visibility("default"))) int g8_gemm_host(g8_host_plan *plan, const void *alpha, const void *hA, size_t lda, const void *hB, size_t ldb,
                                                         const void *beta, void *hC, size_t ldc, void *stream) {
    if (!plan) return G8_STATUS_INVALID_VALUE;
    return run(*reinterpret_cast<HostPlan *>(plan), alpha, hA, lda, hB, ldb, beta, hC, ldc, static_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int g8_host_plan_destroy(g8_host_plan *plan) {
    if (!plan) return 0;
    HostPlan *p = reinterpret_cast<HostPlan *>(plan);
    cudaStreamSynchronize(p->s_h2d), cudaStreamSynchronize(p->s_comp), cudaStreamSynchronize(p->s_d2h);
    destroy(p);
    return 0;
}

} // extern "C"
