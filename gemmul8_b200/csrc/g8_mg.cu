// gemmul8_b200 -- native K-sharded multi-GPU emulated GEMM (include/gemmul8_c.h: g8_mg_comm_*, g8_mg_plan_*, g8_gemm_mg).
//
// New work (the reference is single-GPU).  One process per GPU; rank r owns the K-slab op(A)[:, K_r], op(B)[K_r, :]
// and reconstructs the column slab C[:, n_r].  Everything between the ranks travels over NVLink peer memory (CUDA IPC) through OUR
// OWN kernels -- no NCCL, no MPI, no Python on this path:
//   * bulk exchange : the tcgen05 GEMM epilogue scatters its residue tiles straight into the owners' receive areas
//                     (g8_stage_gemm_scatter: shared memory -> TMA tensor store -> peer HBM, overlapped with the MMAs of the next tile);
//   * small vectors : `mg_allreduce_kernel` -- every rank stores its vector into a mailbox slot of every peer, a flag exchange with
//                     system-scope release / acquire orders the ranks, every rank reduces the slots IN RANK ORDER (so the round-up
//                     sums of fast mode are identical on every rank, which NCCL does not promise);
//   * rank barrier  : the same flag exchange without data (`mg_barrier_kernel`).
// The host only needs a way to pass one 64-byte IPC handle per rank around at start-up (any transport: a pipe, a file, MPI ...).
//
// FP8 backend (run_f8): the piece products are contracted and recombined locally into int16 residues (the single-GPU contract()), the
// owners' column slabs travel as plain peer copies, and the owner sums the shards mod p before the FP8 CRT; K_total <= 2^16.
//
// Exactness: |sum| <= K_total * 2^14 < 2^31 requires K_total = world * k_local <= 2^17.  Accurate mode is bit-identical to the
// single-GPU g8_gemm on the concatenated operands (max and integer sums are order-free); fast mode may differ in a shift on a floor()
// boundary (its sum of squares is reduced per shard, then across shards), and is then identical on all ranks.
#include "g8_internal.cuh"
#include "../../include/gemmul8_c.h"

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

namespace g8 {

constexpr int MG_FLAG_STRIDE = 64;                        // one 64-byte line per source rank
constexpr size_t MG_HEADER   = 2 * G8_MAX_PEERS * MG_FLAG_STRIDE; // flags | error word
constexpr unsigned long long MG_TIMEOUT_CYCLES = 40000000000ull; // ~20 s: a dead peer must not hang the GPU for ever

struct MgDev { // what the kernels need (passed by value)
    int world, rank;
    char *peer[G8_MAX_PEERS]; // every rank's mailbox as seen from this process (peer[rank] = local)
    size_t slot_bytes;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// thread `tid` < world: tell rank `tid` that this rank reached `epoch`, then wait until rank `tid` has reached it too
__device__ __forceinline__ void flag_exchange(const MgDev &d, int tid, uint32_t epoch) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t *>(d.peer[tid] + (size_t)d.rank * MG_FLAG_STRIDE), epoch);
    const uint32_t *mine = reinterpret_cast<const uint32_t *>(d.peer[d.rank] + (size_t)tid * MG_FLAG_STRIDE);
    const unsigned long long t0 = clock64();
    // epochs only grow; "reached or passed" (a peer may already be one collective ahead); wrap-around safe comparison
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
        if (clock64() - t0 > MG_TIMEOUT_CYCLES) {
            *reinterpret_cast<volatile uint32_t *>(d.peer[d.rank] + G8_MAX_PEERS * MG_FLAG_STRIDE) = 1u; // error word
            break;
        }
    }
}

__global__ void mg_barrier_kernel(MgDev d, uint32_t epoch) {
    if ((int)threadIdx.x < d.world) flag_exchange(d, threadIdx.x, epoch);
}

// OP 0: max, 1: sum (rank order 0, 1, ... on EVERY rank -> identical bits everywhere)
template <typename T, int OP> __global__ void __launch_bounds__(1024) mg_allreduce_kernel(MgDev d, const T *__restrict__ src, T *__restrict__ out, int count, uint32_t epoch) {
    const size_t data0 = MG_HEADER + (size_t)(epoch & 1u) * d.world * d.slot_bytes; // double-buffered by the parity of the epoch
    // 1. my vector -> slot [rank] of every peer's mailbox (remote stores over NVLink; own mailbox included)
    for (int o = 0; o < d.world; ++o) {
        T *dst = reinterpret_cast<T *>(d.peer[(d.rank + o) % d.world] + data0 + (size_t)d.rank * d.slot_bytes);
        for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    // 2. all ranks have delivered
    if ((int)threadIdx.x < d.world) flag_exchange(d, threadIdx.x, epoch);
    __syncthreads();
    // 3. reduce the slots of my own mailbox
    const char *base = d.peer[d.rank] + data0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        T acc = reinterpret_cast<const volatile T *>(base)[i];
        for (int o = 1; o < d.world; ++o) {
            const T v = reinterpret_cast<const volatile T *>(base + (size_t)o * d.slot_bytes)[i];
            if constexpr (OP == 0) acc = v > acc ? v : acc;
            else acc = acc + v;
        }
        out[i] = acc;
    }
}

struct MgComm {
    int world = 0, rank = 0, device = 0;
    size_t slot_bytes = 0, bytes = 0;
    char *local = nullptr;
    char *peer[G8_MAX_PEERS] = {};
    bool connected = false;
    uint32_t epoch = 0; // every rank issues the same sequence of collectives
    MgDev dev() const {
        MgDev d{};
        d.world = world, d.rank = rank, d.slot_bytes = slot_bytes;
        for (int o = 0; o < world; ++o) d.peer[o] = peer[o];
        return d;
    }
};

static int comm_barrier(MgComm &c, cudaStream_t st) {
    mg_barrier_kernel<<<1, 32, 0, st>>>(c.dev(), ++c.epoch);
    return (int)cudaGetLastError();
}
template <typename T, int OP> static int comm_allreduce(MgComm &c, const T *src, T *out, size_t count, cudaStream_t st) {
    if (count * sizeof(T) > c.slot_bytes) return G8_STATUS_NOT_SUPPORTED;
    if (count == 0) return 0;
    mg_allreduce_kernel<T, OP><<<1, 1024, 0, st>>>(c.dev(), src, out, (int)count, ++c.epoch);
    return (int)cudaGetLastError();
}

struct MgPlan {
    MgComm *comm = nullptr;
    int dtype = F64, opA = OP_N, opB = OP_N, fast = 0, backend = INT8;
    size_t nmat = 0;               // planes per plane set (INT8: N; FP8: 2 or 3 e4m3 pieces per modulus)
    int8_t *scratch = nullptr;     // FP8: per-product residues of a batch of moduli (contract())
    size_t scratch_bytes = 0;
    int8_t *C_loc = nullptr;       // FP8: this rank's partial C_mid for ALL n columns, int16 [modulus][n][m_pad] (complex: {re, im})
    size_t m = 0, n = 0, k = 0, m_pad = 0, k_pad = 0, n_pad = 0, nc = 0, sizeA = 0, sizeB = 0;
    unsigned N = 0;
    bool cplx = false;
    size_t sets = 1, bsets = 1; // plane sets of the residues (Re, Im, Re+Im: 3 for complex) / of the bound planes (|Re|, |Im|: 2)
    int8_t *A_lo = nullptr, *B_lo = nullptr, *C_mid = nullptr; // C_mid: complex only (owner-side 3M recombination before the CRT)
    int16_t *sftA = nullptr, *sftB = nullptr;
    double *stat = nullptr;   // amax[m + n] | sumsq[m + n] | reduced copies
    int32_t *maxes = nullptr; // rowmax[m_pad] | colmax[n_pad] (input of the MAX all-reduce) | reduced copy
    char *recv = nullptr;     // my receive area: [src rank][modulus][col in slab][row] int8, then (accurate mode) the gathered bound
                              // planes: A-bar [src rank][m_pad][k_pad] and the B-bar columns of MY slab [src rank][nc][k_pad]
    size_t recv_bytes = 0, abar_off = 0, bbar_off = 0;
    char *peer_recv[G8_MAX_PEERS] = {};
    // helper stream for the B side of the preprocessing (forked from / joined into the caller's stream inside every call)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

#define G8_TRY(x)                              \
    do {                                       \
        const int _e = (int)(x);               \
        if (_e != 0) return _e;                \
    } while (0)

static void plan_free(MgPlan *p) {
    if (!p) return;
    for (void *q : {(void *)p->A_lo, (void *)p->B_lo, (void *)p->C_mid, (void *)p->sftA, (void *)p->sftB, (void *)p->stat, (void *)p->maxes, (void *)p->scratch,
                    (void *)p->C_loc})
        if (q) cudaFree(q);
    if (p->comm)
        for (int o = 0; o < p->comm->world; ++o)
            if (o != p->comm->rank && p->peer_recv[o]) cudaIpcCloseMemHandle(p->peer_recv[o]);
    if (p->recv) cudaFree(p->recv);
    if (p->side) cudaStreamDestroy(p->side);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    delete p;
}

// all ranks contribute `bytes` (<= slot) of host data, every rank gets all of them in rank order (start-up only: synchronises)
static int comm_exchange_host(MgComm &c, const void *mine, size_t bytes, void *all) {
    if (bytes > c.slot_bytes || bytes % 4) return G8_STATUS_INVALID_VALUE;
    uint32_t *d_in = nullptr, *d_out = nullptr;
    const size_t words = bytes / 4, W = (size_t)c.world;
    G8_TRY(cudaMalloc(&d_in, W * bytes));
    G8_TRY(cudaMalloc(&d_out, W * bytes));
    cudaMemset(d_in, 0, W * bytes);
    cudaMemcpy(reinterpret_cast<char *>(d_in) + (size_t)c.rank * bytes, mine, bytes, cudaMemcpyHostToDevice);
    int e = 0;
    // gather == MAX all-reduce of a vector that is zero outside this rank's section (handles are arbitrary bits: reduce as uint32)
    for (size_t off = 0; off < W * words && !e; off += c.slot_bytes / 4) {
        const size_t cnt = std::min(c.slot_bytes / 4, W * words - off);
        e = comm_allreduce<uint32_t, 0>(c, d_in + off, d_out + off, cnt, nullptr);
    }
    if (!e) e = (int)cudaMemcpy(all, d_out, W * bytes, cudaMemcpyDeviceToHost);
    cudaFree(d_in), cudaFree(d_out);
    return e;
}

static int plan_create(MgPlan **out, MgComm *c, int dtype, int backend, int opA, int opB, size_t m, size_t n, size_t k_local, unsigned N, int fast) {
    if (backend != INT8 && backend != FP8) return G8_STATUS_INVALID_VALUE;
    if (!out || !c || !c->connected || dtype < F32 || dtype > C64 || opA < 0 || opA > 2 || opB < 0 || opB > 2) return G8_STATUS_INVALID_VALUE;
    if (N < 2 || N > G8_MAX_MODULI || m == 0 || n == 0 || k_local == 0) return G8_STATUS_INVALID_VALUE;
    const size_t W = (size_t)c->world;
    if (n % W || (n / W) % 256) return G8_STATUS_INVALID_VALUE;              // the scatter hands whole 256-column tiles to one owner
    if (W * k_local > (size_t(1) << 17)) return G8_STATUS_INVALID_VALUE;       // INT32 accumulation bound over the TOTAL K
    if (backend == FP8 && W * k_local > (size_t(1) << 16)) return G8_STATUS_INVALID_VALUE; // binary32 accumulation of the piece products (g8_gemm)
    if (!device_supported_cached()) return G8_STATUS_NO_DEVICE_CODE;
    MgPlan *p = new (std::nothrow) MgPlan();
    if (!p) return (int)cudaErrorMemoryAllocation;
    p->backend = backend, p->nmat = num_planes(backend, N);
    p->comm = c, p->dtype = dtype, p->opA = opA, p->opB = opB, p->fast = fast, p->m = m, p->n = n, p->k = k_local, p->N = N;
    p->m_pad = pad256(m), p->k_pad = pad256(k_local), p->n_pad = pad256(n), p->nc = n / W;
    p->sizeA = p->k_pad * p->m_pad, p->sizeB = p->k_pad * n;
    p->cplx = dtype >= C32, p->sets = p->cplx ? 3 : 1, p->bsets = p->cplx ? 2 : 1;
    if ((m + n) * sizeof(double) > c->slot_bytes || (p->m_pad + p->n_pad) * sizeof(int32_t) > c->slot_bytes) {
        delete p;
        return G8_STATUS_NOT_SUPPORTED; // mailbox slots too small for this problem: create the communicator with a larger max_vector_bytes
    }
    auto fail = [&](int code) {
        plan_free(p);
        return code;
    };
#define G8_ALLOC(ptr, bytes) \
    if (cudaMalloc(reinterpret_cast<void **>(&(ptr)), (bytes)) != cudaSuccess) return fail((int)cudaErrorMemoryAllocation)
    const size_t mid = (backend == FP8 ? 2 : 1) * (p->cplx ? 2 : 1); // bytes per C_mid element
    G8_ALLOC(p->A_lo, p->sizeA * p->nmat * p->sets);
    G8_ALLOC(p->B_lo, p->sizeB * p->nmat * p->sets);
    if (p->cplx || backend == FP8) G8_ALLOC(p->C_mid, mid * (size_t)N * p->nc * p->m_pad);
    if (backend == FP8) {
        G8_ALLOC(p->C_loc, mid * (size_t)N * n * p->m_pad);
        const size_t per_mod = (p->cplx ? 9 : 3) * sizeof(int16_t) * p->m_pad * n; // contract(): per-product residues of one modulus
        p->scratch_bytes     = per_mod * std::min<size_t>(N, 4);
        G8_ALLOC(p->scratch, p->scratch_bytes);
    }
    G8_ALLOC(p->sftA, sizeof(int16_t) * p->m_pad);
    G8_ALLOC(p->sftB, sizeof(int16_t) * p->n_pad);
    G8_ALLOC(p->stat, sizeof(double) * 4 * (m + n));
    G8_ALLOC(p->maxes, sizeof(int32_t) * 2 * (p->m_pad + p->n_pad));
#undef G8_ALLOC
    cudaMemset(p->sftA, 0, sizeof(int16_t) * p->m_pad), cudaMemset(p->sftB, 0, sizeof(int16_t) * p->n_pad);
    if (cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming) != cudaSuccess)
        return fail((int)cudaErrorUnknown);
    // INT8: W shards x (N moduli [x 3 products]) x nc columns x m_pad rows (int8); FP8: W shards x N moduli x nc x m_pad C_mid elements
    const size_t per = backend == FP8 ? mid * (size_t)N * n * p->m_pad : p->sets * (size_t)N * n * p->m_pad;
    p->abar_off   = per;
    p->bbar_off   = per + W * p->bsets * p->sizeA;
    p->recv_bytes = per + (fast ? 0 : W * p->bsets * (p->sizeA + p->nc * p->k_pad));
    unsigned char handle[64];
    if (int e = g8_peer_alloc(p->recv_bytes, reinterpret_cast<void **>(&p->recv), handle)) return fail(e);
    std::vector<unsigned char> all(64 * W);
    if (int e = comm_exchange_host(*c, handle, 64, all.data())) return fail(e);
    for (size_t o = 0; o < W; ++o) {
        if ((int)o == c->rank) {
            p->peer_recv[o] = p->recv;
            continue;
        }
        void *q = nullptr;
        if (int e = g8_peer_open(all.data() + 64 * o, &q)) return fail(e);
        p->peer_recv[o] = static_cast<char *>(q);
    }
    *out = p;
    return 0;
}

static int run_f8(MgPlan &p, const void *alpha, const void *A, size_t lda, const void *B, size_t ldb, const void *beta, void *C, size_t ldc, cudaStream_t st);

static int run(MgPlan &p, const void *alpha, const void *A, size_t lda, const void *B, size_t ldb, const void *beta, void *C, size_t ldc, cudaStream_t st) {
    if (p.backend == FP8) return run_f8(p, alpha, A, lda, B, ldb, beta, C, ldc, st);
    MgComm &c = *p.comm;
    const size_t m = p.m, n = p.n, k = p.k, W = (size_t)c.world, nc = p.nc, mp = p.m_pad;
    const unsigned N = p.N;
    if (!alpha || !beta || !A || !B || !C) return G8_STATUS_INVALID_VALUE;
    double *amax = p.stat, *ss = p.stat + (m + n), *amax_r = p.stat + 2 * (m + n), *ss_r = p.stat + 3 * (m + n);
    void *pst = st;

    // The A side and the B side are independent until they meet in a GEMM: B-side kernels run on the plan's helper stream
    // (fork: it waits for `st`; join: `st` waits for it), which also lets the A-bar peer copies overlap the B-bar kernel.
    void *psd = p.side;
    auto fork = [&]() -> int {
        G8_TRY(cudaEventRecord(p.ev_fork, st));
        return (int)cudaStreamWaitEvent(p.side, p.ev_fork, 0);
    };
    auto join = [&]() -> int {
        G8_TRY(cudaEventRecord(p.ev_join, p.side));
        return (int)cudaStreamWaitEvent(st, p.ev_join, 0);
    };

    // ---- shifts from GLOBAL row statistics ----
    G8_TRY(fork());
    G8_TRY(g8_stage_stats(p.dtype, 1, p.opA, m, k, A, lda, amax, ss, pst));
    G8_TRY(g8_stage_stats(p.dtype, 0, p.opB, n, k, B, ldb, amax + m, ss + m, psd));
    G8_TRY(join());
    G8_TRY((comm_allreduce<double, 0>(c, amax, amax_r, m + n, st)));
    if (p.fast) {
        G8_TRY((comm_allreduce<double, 1>(c, ss, ss_r, m + n, st)));
        G8_TRY(fork());
        G8_TRY(g8_stage_shift_from_stats(amax_r, ss_r, m, N, 0, p.sftA, pst));
        G8_TRY(g8_stage_shift_from_stats(amax_r + m, ss_r + m, n, N, 0, p.sftB, psd));
    } else {
        // accurate: s0 from the global max, bound planes (aliasing plane 0).  The int8 bound planes are exchanged instead of INT32
        // partial products (4x .. 16x fewer bytes): every rank receives all K-slabs of A-bar and the K-slabs of ITS column slab of B-bar
        // (plain peer copies on the copy engines), multiplies them over the FULL K with the maxima fused into the GEMM epilogue, then
        // one MAX all-reduce of [row maxima | column maxima] gives everybody the final shifts.
        G8_TRY(fork());
        G8_TRY(g8_stage_shift_from_stats(amax_r, nullptr, m, N, 1, p.sftA, pst));
        G8_TRY(g8_stage_split(p.dtype, 1, p.opA, m, k, A, lda, N, 3, p.sftA, p.A_lo, p.sizeA, N, pst));
        G8_TRY(g8_stage_shift_from_stats(amax_r + m, nullptr, n, N, 1, p.sftB, psd));
        G8_TRY(g8_stage_split(p.dtype, 0, p.opB, n, k, B, ldb, N, 3, p.sftB, p.B_lo, p.sizeB, N, psd));
        // gathered layout: A-bar [slab][|Re| (,|Im|)][m_pad][k_pad], B-bar [slab][|Re| (,|Im|)][nc][k_pad]
        const size_t bslab = nc * p.k_pad, bs = p.bsets;
        for (size_t j = 0; j < W; ++j) { // A-bar travels while the B-bar kernel still runs on the helper stream
            const size_t o = ((size_t)c.rank + j) % W; // staggered targets
            G8_TRY(cudaMemcpyAsync(p.peer_recv[o] + p.abar_off + (size_t)c.rank * bs * p.sizeA, p.A_lo, bs * p.sizeA, cudaMemcpyDefault, st));
        }
        G8_TRY(join());
        for (size_t j = 0; j < W; ++j) {
            const size_t o = ((size_t)c.rank + j) % W;
            for (size_t g = 0; g < bs; ++g)
                G8_TRY(cudaMemcpyAsync(p.peer_recv[o] + p.bbar_off + ((size_t)c.rank * bs + g) * bslab, p.B_lo + g * p.sizeB + o * bslab, bslab, cudaMemcpyDefault, st));
        }
        G8_TRY(comm_barrier(c, st));
        int32_t *mx = p.maxes, *mx_r = p.maxes + (p.m_pad + p.n_pad);
        G8_TRY(cudaMemsetAsync(mx, 0, sizeof(int32_t) * (p.m_pad + p.n_pad), st));
        if (!p.cplx) {
            G8_TRY(g8_stage_gemm_bound_chain(reinterpret_cast<const int8_t *>(p.recv + p.abar_off), p.sizeA, reinterpret_cast<const int8_t *>(p.recv + p.bbar_off), bslab,
                                             m, nc, p.k_pad, (int)W, mx, mx + p.m_pad + (size_t)c.rank * nc, pst));
        } else {
            GemmArgs g{};
            g.A = reinterpret_cast<const int8_t *>(p.recv + p.abar_off), g.B = reinterpret_cast<const int8_t *>(p.recv + p.bbar_off);
            g.strideA = p.sizeA, g.strideB = bslab, g.m = m, g.n = nc, g.m_pad = mp, g.k_pad = p.k_pad;
            g.num_units = 1, g.first_modulus = 0, g.epi = EPI_BOUND_MAX_CPLX, g.kchain = (int)W, g.k_true = (int)p.k_pad;
            g.groupA[0] = 0, g.groupA[1] = 1, g.groupA[2] = 2, g.groupB[0] = 0, g.groupB[1] = 1, g.groupB[2] = 2;
            g.ldc = mp, g.rowmax = mx, g.colmax = mx + p.m_pad + (size_t)c.rank * nc;
            G8_TRY(launch_gemm_tc(g, st));
        }
        G8_TRY((comm_allreduce<int32_t, 0>(c, mx, mx_r, p.m_pad + p.n_pad, st)));
        G8_TRY(fork());
        G8_TRY(g8_stage_finalize_shift(p.sftA, mx_r, m, N, pst));
        G8_TRY(g8_stage_finalize_shift(p.sftB, mx_r + p.m_pad, n, N, psd));
    }

    // ---- local split with the global shifts (A on `st`, B on the helper stream), contraction fused with the exchange, owner-side sum + CRT ----
    G8_TRY(g8_stage_split(p.dtype, 1, p.opA, m, k, A, lda, N, 0, p.sftA, p.A_lo, p.sizeA, N, pst));
    G8_TRY(g8_stage_split(p.dtype, 0, p.opB, n, k, B, ldb, N, 0, p.sftB, p.B_lo, p.sizeB, N, psd));
    G8_TRY(join());
    const size_t units = p.sets * (size_t)N; // real: one product per modulus; complex: the three 3M products per modulus
    {
        GemmArgs g{};
        g.A = p.A_lo, g.B = p.B_lo, g.strideA = p.sizeA, g.strideB = p.sizeB, g.m = m, g.n = n, g.m_pad = mp, g.k_pad = p.k_pad;
        g.num_units = (int)units, g.first_modulus = 0, g.epi = EPI_MOD_I8, g.k_true = (int)p.k_pad;
        g.prods = p.cplx ? 3 : 1;
        for (int i = 0; i < 3; ++i) g.groupA[i] = g.groupB[i] = p.cplx ? i * (int)N : 0; // plane sets Re / Im / Re+Im, N planes each
        g.out = nullptr, g.out_stride = nc * mp, g.ldc = mp;
        for (size_t o = 0; o < W; ++o) g.peer_out[o] = p.peer_recv[o] + (size_t)c.rank * units * nc * mp;
        g.owner_cols = nc, g.rank = c.rank, g.world = (int)W;
        G8_TRY(launch_gemm_tc(g, st));
    }
    G8_TRY(comm_barrier(c, st)); // every rank's tiles have landed (kernel completion + system-scope release / acquire)
    if (!p.cplx) {
        G8_TRY(g8_stage_crt_parts(p.dtype, p.recv, (int)W, (size_t)N * nc * mp, mp, nc * mp, m, nc, N, C, ldc, p.sftA, p.sftB + (size_t)c.rank * nc, alpha, beta, pst));
    } else {
        // owner side, complex: sum the shards per 3M product, reduce, recombine {Re, Im} mod p, then the ordinary complex CRT
        launch_i8_cplx_combine_parts(reinterpret_cast<const int8_t *>(p.recv), (int)W, units * nc * mp, nc * mp, (int)N, 0, p.C_mid, nc * mp, st);
        G8_TRY(g8_stage_crt(p.dtype, p.C_mid, mp, nc * mp, m, nc, N, C, ldc, p.sftA, p.sftB + (size_t)c.rank * nc, alpha, beta, pst));
    }
    // (the next call's first all-reduce orders every rank's CRT before anybody's next scatter into the receive areas)
    return (int)cudaPeekAtLastError();
}


// ---- FP8 backend ----
// Same shift protocol as the INT8 path (global statistics; accurate mode: gathered e4m3 bound planes, bound GEMM chained over the K-slabs in
// rank order -- the binary32 accumulation then performs the same sequence of MMAs as the single-GPU call whenever k_local is a multiple of
// 32, the extra zero padding of each slab adding exact zeros).  Contraction: the single-GPU contract() over all n columns into a local
// int16 C_mid; exchange: one strided peer copy per owner; owner: shard sum mod p (i16_sum_parts) + the FP8 CRT.
static int run_f8(MgPlan &p, const void *alpha, const void *A, size_t lda, const void *B, size_t ldb, const void *beta, void *C, size_t ldc, cudaStream_t st) {
    MgComm &c = *p.comm;
    const size_t m = p.m, n = p.n, k = p.k, W = (size_t)c.world, nc = p.nc, mp = p.m_pad;
    const unsigned N = p.N;
    if (!alpha || !beta || !A || !B || !C) return G8_STATUS_INVALID_VALUE;
    double *amax = p.stat, *ss = p.stat + (m + n), *amax_r = p.stat + 2 * (m + n), *ss_r = p.stat + 3 * (m + n);
    auto fork = [&]() -> int {
        G8_TRY(cudaEventRecord(p.ev_fork, st));
        return (int)cudaStreamWaitEvent(p.side, p.ev_fork, 0);
    };
    auto join = [&]() -> int {
        G8_TRY(cudaEventRecord(p.ev_join, p.side));
        return (int)cudaStreamWaitEvent(st, p.ev_join, 0);
    };
    const SplitArgs sa = make_split_args(1, p.opA, m, k, A, lda, N, p.sftA, p.A_lo, p.sizeA, p.nmat, FP8);
    const SplitArgs sb = make_split_args(0, p.opB, n, k, B, ldb, N, p.sftB, p.B_lo, p.sizeB, p.nmat, FP8);

    G8_TRY(fork());
    launch_stats(sa, p.dtype, amax, ss, st);
    launch_stats(sb, p.dtype, amax + m, ss + m, p.side);
    G8_TRY(join());
    G8_TRY((comm_allreduce<double, 0>(c, amax, amax_r, m + n, st)));
    if (p.fast) {
        G8_TRY((comm_allreduce<double, 1>(c, ss, ss_r, m + n, st)));
        G8_TRY(fork());
        launch_shift_from_stats(amax_r, ss_r, m, (int)N, 0, p.sftA, st, FP8);
        launch_shift_from_stats(amax_r + m, ss_r + m, n, (int)N, 0, p.sftB, p.side, FP8);
    } else {
        G8_TRY(fork());
        SplitArgs ea = sa, eb = sb; // bound planes alias plane 0 (,1) of the residue planes
        for (int g = 0; g < 3; ++g) ea.planes[g] = p.A_lo + g * p.sizeA, eb.planes[g] = p.B_lo + g * p.sizeB;
        launch_shift_from_stats(amax_r, nullptr, m, (int)N, 1, p.sftA, st, FP8);
        launch_split(ea, p.dtype, 3, st);
        launch_shift_from_stats(amax_r + m, nullptr, n, (int)N, 1, p.sftB, p.side, FP8);
        launch_split(eb, p.dtype, 3, p.side);
        const size_t bslab = nc * p.k_pad, bs = p.bsets;
        for (size_t j = 0; j < W; ++j) {
            const size_t o = ((size_t)c.rank + j) % W;
            G8_TRY(cudaMemcpyAsync(p.peer_recv[o] + p.abar_off + (size_t)c.rank * bs * p.sizeA, p.A_lo, bs * p.sizeA, cudaMemcpyDefault, st));
        }
        G8_TRY(join());
        for (size_t j = 0; j < W; ++j) {
            const size_t o = ((size_t)c.rank + j) % W;
            for (size_t g = 0; g < bs; ++g)
                G8_TRY(cudaMemcpyAsync(p.peer_recv[o] + p.bbar_off + ((size_t)c.rank * bs + g) * bslab, p.B_lo + g * p.sizeB + o * bslab, bslab, cudaMemcpyDefault, st));
        }
        G8_TRY(comm_barrier(c, st));
        int32_t *mx = p.maxes, *mx_r = p.maxes + (p.m_pad + p.n_pad);
        G8_TRY(cudaMemsetAsync(mx, 0, sizeof(int32_t) * (p.m_pad + p.n_pad), st));
        GemmArgs g{};
        g.A = reinterpret_cast<const int8_t *>(p.recv + p.abar_off), g.B = reinterpret_cast<const int8_t *>(p.recv + p.bbar_off);
        g.strideA = p.sizeA, g.strideB = bslab, g.m = m, g.n = nc, g.m_pad = mp, g.k_pad = p.k_pad;
        g.num_units = 1, g.first_modulus = 0, g.epi = p.cplx ? EPI_F8_BOUND_CPLX : EPI_F8_BOUND, g.kchain = (int)W;
        g.k_true = (int)(W * k); // the inflation factor (k + 1) * 2^-24 of the un-sharded product
        g.groupA[0] = 0, g.groupA[1] = 1, g.groupA[2] = 2, g.groupB[0] = 0, g.groupB[1] = 1, g.groupB[2] = 2;
        if (!p.cplx) g.groupA[1] = g.groupA[2] = g.groupB[1] = g.groupB[2] = 0;
        g.ldc = mp, g.rowmax = mx, g.colmax = mx + p.m_pad + (size_t)c.rank * nc;
        G8_TRY(launch_gemm_tc(g, st));
        // the maxima are non-negative floats stored by their bit pattern: the integer MAX all-reduce orders them correctly
        G8_TRY((comm_allreduce<int32_t, 0>(c, mx, mx_r, p.m_pad + p.n_pad, st)));
        G8_TRY(fork());
        launch_finalize_accu_shift(p.sftA, mx_r, m, (int)N, st, FP8);
        launch_finalize_accu_shift(p.sftB, mx_r + p.m_pad, n, (int)N, p.side, FP8);
    }
    launch_split(sa, p.dtype, 0, st);
    launch_split(sb, p.dtype, 0, p.side);
    G8_TRY(join());

    // ---- contraction of my K-slab for all n columns, then every owner's column slab of every modulus travels as one strided peer copy ----
    const size_t esz = sizeof(int16_t) * (p.cplx ? 2 : 1), slab = nc * mp * esz;
    ContractArgs ca{};
    ca.cplx = p.cplx, ca.backend = FP8, ca.N = N, ca.m = m, ca.ncols = n, ca.m_pad = mp, ca.k_pad = p.k_pad;
    ca.A_lo = p.A_lo, ca.B_lo = p.B_lo, ca.sizeA = p.sizeA, ca.sizeB = p.sizeB, ca.set_planes = p.nmat;
    ca.C_mid = p.C_loc, ca.mid_plane_stride = mp * n, ca.scratch = p.scratch, ca.scratch_avail = p.scratch_bytes;
    G8_TRY(contract(ca, st));
    for (size_t j = 0; j < W; ++j) {
        const size_t o = ((size_t)c.rank + j) % W;
        char *dst = p.peer_recv[o] + (size_t)c.rank * N * slab;
        if (W * slab < (size_t(1) << 31)) { // pitch limit of the 2-D copy
            G8_TRY(cudaMemcpy2DAsync(dst, slab, p.C_loc + o * slab, W * slab, slab, N, cudaMemcpyDefault, st));
        } else {
            for (size_t u = 0; u < N; ++u) G8_TRY(cudaMemcpyAsync(dst + u * slab, p.C_loc + (u * W + o) * slab, slab, cudaMemcpyDefault, st));
        }
    }
    G8_TRY(comm_barrier(c, st));
    const size_t unit_elems = nc * mp * (p.cplx ? 2 : 1); // int16 values per modulus of my slab
    launch_i16_sum_parts(reinterpret_cast<const int16_t *>(p.recv), (int)W, (size_t)N * unit_elems, unit_elems, (int)N, 0, reinterpret_cast<int16_t *>(p.C_mid),
                         unit_elems, st);
    CrtArgs cr{};
    cr.C_mid = p.C_mid, cr.ldmid = mp, cr.plane_stride = nc * mp, cr.m = m, cr.n = nc, cr.num_moduli = (int)N;
    cr.C = C, cr.ldc = ldc, cr.sftA = p.sftA, cr.sftB = p.sftB + (size_t)c.rank * nc, cr.alpha = alpha, cr.beta = beta, cr.backend = FP8;
    G8_TRY(launch_crt(cr, p.dtype, st));
    return (int)cudaPeekAtLastError();
}

} // namespace g8

using namespace g8;

extern "C" {

__attribute__((visibility("default"))) int g8_mg_comm_create(g8_mg_comm **comm, int world, int rank, size_t max_vector_bytes, void *handle64) {
    if (!comm || !handle64 || world < 1 || world > G8_MAX_PEERS || rank < 0 || rank >= world) return G8_STATUS_INVALID_VALUE;
    MgComm *c = new (std::nothrow) MgComm();
    if (!c) return (int)cudaErrorMemoryAllocation;
    c->world = world, c->rank = rank;
    cudaGetDevice(&c->device);
    c->slot_bytes = std::max<size_t>((max_vector_bytes + 255) / 256 * 256, 4096);
    c->bytes      = MG_HEADER + 2 * (size_t)world * c->slot_bytes;
    if (int e = g8_peer_alloc(c->bytes, reinterpret_cast<void **>(&c->local), handle64)) {
        delete c;
        return e;
    }
    cudaMemset(c->local, 0, c->bytes);
    cudaDeviceSynchronize();
    c->peer[rank] = c->local;
    *comm = reinterpret_cast<g8_mg_comm *>(c);
    return 0;
}

__attribute__((visibility("default"))) int g8_mg_comm_connect(g8_mg_comm *comm, const void *handles) {
    MgComm *c = reinterpret_cast<MgComm *>(comm);
    if (!c || !handles || c->connected) return G8_STATUS_INVALID_VALUE;
    for (int o = 0; o < c->world; ++o) {
        if (o == c->rank) continue;
        void *q = nullptr;
        if (int e = g8_peer_open(static_cast<const char *>(handles) + 64 * (size_t)o, &q)) return e;
        c->peer[o] = static_cast<char *>(q);
    }
    c->connected = true;
    return 0;
}

__attribute__((visibility("default"))) int g8_mg_comm_barrier(g8_mg_comm *comm, void *stream) {
    MgComm *c = reinterpret_cast<MgComm *>(comm);
    if (!c || !c->connected) return G8_STATUS_INVALID_VALUE;
    return comm_barrier(*c, static_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int g8_mg_comm_status(g8_mg_comm *comm) {
    MgComm *c = reinterpret_cast<MgComm *>(comm);
    if (!c) return G8_STATUS_INVALID_VALUE;
    uint32_t err = 0;
    if (cudaMemcpy(&err, c->local + G8_MAX_PEERS * MG_FLAG_STRIDE, sizeof(err), cudaMemcpyDeviceToHost) != cudaSuccess) return (int)cudaGetLastError();
    return err ? (int)cudaErrorTimeout : 0;
}

__attribute__((visibility("default"))) int g8_mg_comm_destroy(g8_mg_comm *comm) {
    MgComm *c = reinterpret_cast<MgComm *>(comm);
    if (!c) return 0;
    cudaDeviceSynchronize();
    for (int o = 0; o < c->world; ++o)
        if (o != c->rank && c->peer[o]) cudaIpcCloseMemHandle(c->peer[o]);
    if (c->local) cudaFree(c->local);
    delete c;
    return 0;
}

__attribute__((visibility("default"))) int g8_mg_plan_create(g8_mg_plan **plan, g8_mg_comm *comm, int dtype, int op_A, int op_B, size_t m, size_t n, size_t k_local,
                                                              unsigned num_moduli, int fastmode) {
    return g8_mg_plan_create_backend(plan, comm, dtype, G8_BACKEND_INT8, op_A, op_B, m, n, k_local, num_moduli, fastmode);
}

__attribute__((visibility("default"))) int g8_mg_plan_create_backend(g8_mg_plan **plan, g8_mg_comm *comm, int dtype, int backend, int op_A, int op_B, size_t m, size_t n,
                                                                      size_t k_local, unsigned num_moduli, int fastmode) {
    MgPlan *p = nullptr;
    const int e = plan_create(&p, reinterpret_cast<MgComm *>(comm), dtype, backend, op_A, op_B, m, n, k_local, num_moduli, fastmode != 0);
    if (e == 0) *plan = reinterpret_cast<g8_mg_plan *>(p);
    return e;
}

__attribute__((visibility("default"))) int g8_gemm_mg(g8_mg_plan *plan, const void *alpha, const void *A_local, size_t lda, const void *B_local, size_t ldb,
                                                       const void *beta, void *C_slab, size_t ldc, void *stream) {
    if (!plan) return G8_STATUS_INVALID_VALUE;
    return run(*reinterpret_cast<MgPlan *>(plan), alpha, A_local, lda, B_local, ldb, beta, C_slab, ldc, static_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int g8_mg_plan_destroy(g8_mg_plan *plan) {
    MgPlan *p = reinterpret_cast<MgPlan *>(plan);
    if (!p) return 0;
    cudaDeviceSynchronize();
    if (p->comm && p->comm->connected) { // nobody may still have our receive area mapped / in use when it is freed
        comm_barrier(*p->comm, nullptr);
        cudaDeviceSynchronize();
    }
    plan_free(p);
    return 0;
}

} // extern "C"
