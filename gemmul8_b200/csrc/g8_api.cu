// gemmul8_b200 -- C-ABI orchestrator (include/gemmul8_c.h): workspace carving + the three-stage pipeline.
//
// Mirrors the control flow of reference real::gemm / complex::gemm (gemmul8_real.hpp:53-211,
// gemmul8_complex.hpp:53-226) with these deliberate differences:
//   * one persistent tcgen05 launch covers all moduli, mod-p is fused into its epilogue (no C_hi, no conv_hi2mid);
//   * accurate mode never materialises the bound product: its GEMM epilogue reduces row/col maxima directly;
//   * no host synchronisation unless the caller asks for the phase timings;
//   * constants live in static __constant__ memory (no per-call cudaMemcpyToSymbol, table.hpp:854-862).
// The workspace layout (A_lo | sftA | B_lo | sftB | C_mid | scratch) is byte-compatible with the reference's.
#include "g8_internal.cuh"
#include "../../include/gemmul8_c.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace g8 {

struct Sizes {
    size_t k_pad, m_pad, n_pad, sizeA, sizeB, sizeC;
    unsigned num_mat;
    size_t low, mid, hi;      // bytes per element of the low / mid / hi types
    unsigned cplx3, num_C_hi; // 3 for complex (three plane groups), C_hi planes
    size_t totalA, totalB, totalC;
};

static unsigned num_mat(int backend, unsigned n) { // table.hpp:69-75
    if (backend == INT8) return n;
    return n <= 6 ? 2 * n : 12 + 3 * (n - 6);
}
unsigned num_planes(int backend, unsigned num_moduli) { return num_mat(backend, num_moduli); }

static Sizes sizes(bool cplx, int backend, size_t m, size_t n, size_t k, unsigned N, bool enA, bool enB) {
    Sizes s{};
    s.k_pad = pad256(k), s.m_pad = pad256(m), s.n_pad = pad256(n);
    s.sizeA = s.k_pad * s.m_pad, s.sizeB = s.k_pad * n, s.sizeC = s.m_pad * n;
    s.num_mat = num_mat(backend, N);
    s.low = 1, s.mid = (backend == INT8 ? 1 : 2) * (cplx ? 2 : 1), s.hi = 4;
    s.cplx3    = cplx ? 3 : 1;
    s.num_C_hi = (backend == INT8 ? 1 : 3) * s.cplx3;
    const size_t lwork = size_t(1) << 25; // 32 MiB reserved by the reference for cuBLASLt (gemmul8_real.hpp:32)
    s.totalA = 255 + s.low * s.sizeA * (s.num_mat + (enA ? 1 : 0)) * s.cplx3 + sizeof(int16_t) * s.m_pad;
    s.totalB = 255 + s.low * s.sizeB * (s.num_mat + (enB ? 1 : 0)) * s.cplx3 + sizeof(int16_t) * s.n_pad;
    s.totalC = 255 + s.mid * s.sizeC * (N - 1) + std::max(lwork, s.mid * s.sizeC) + s.hi * s.sizeC * s.num_C_hi;
    return s;
}

struct PhaseTimer {
    cudaEvent_t ev[5]{};
    cudaStream_t st;
    bool on;
    PhaseTimer(bool enable, cudaStream_t s) : st(s), on(enable) {
        if (on)
            for (auto &e : ev) cudaEventCreate(&e);
    }
    void mark(int i) {
        if (on) cudaEventRecord(ev[i], st);
    }
    void finish(double *ns) {
        if (!on) return;
        cudaEventSynchronize(ev[4]);
        float ms;
        cudaEventElapsedTime(&ms, ev[0], ev[1]); ns[0] = ms * 1e6;
        cudaEventElapsedTime(&ms, ev[1], ev[2]); ns[1] = ms * 1e6;
        ns[2] = 0.0; // requantisation is fused into the GEMM epilogue
        cudaEventElapsedTime(&ms, ev[2], ev[4]); ns[3] = ms * 1e6;
    }
    ~PhaseTimer() { // also on the early error returns of gemm_impl
        if (on)
            for (auto &e : ev) cudaEventDestroy(e);
    }
    PhaseTimer(const PhaseTimer &) = delete;
    PhaseTimer &operator=(const PhaseTimer &) = delete;
};

static bool device_ok() {
    static std::atomic<int> ok[64]; // per device: 0 unknown, 1 yes, 2 no (racing first calls compute the same value)
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    int v = ok[dev & 63].load(std::memory_order_relaxed);
    if (!v) {
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
        v = (major == 10 && minor == 0) ? 1 : 2;
        ok[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v == 1;
}

SplitArgs make_split_args(int is_A, int op, size_t rows, size_t k, const void *X, size_t ld, unsigned N, int16_t *sft, int8_t *base,
                          size_t plane_stride, size_t group_stride_planes, int backend);
static SplitArgs split_args(int is_A, int op, size_t rows, size_t k, const void *X, size_t ld, unsigned N, int16_t *sft,
                            int8_t *base, size_t plane_stride, size_t group_stride_planes, int backend = INT8) {
    SplitArgs a{};
    a.backend = backend;
    a.X = X, a.ld = ld, a.rows = rows, a.inner = k, a.k_pad = pad256(k), a.sft = sft;
    for (int g = 0; g < 3; ++g) a.planes[g] = base + g * group_stride_planes * plane_stride;
    a.plane_stride = plane_stride;
    a.num_moduli   = (int)N;
    // A: op N -> element (r,l) at A[l*lda + r] (row-strided); op T/C -> row r contiguous.  B: the other way round.
    a.row_contig = is_A ? (op != OP_N) : (op == OP_N);
    a.conj       = (op == OP_C);
    return a;
}

// ---- stage 2 for `ncols` columns of op(B): the low-precision products of all moduli + requantisation into C_mid ----
// Shared by the monolithic call (ncols = n) and the host-buffer pipeline (one column chunk at a time: B_lo / C_mid point at the
// chunk's first column, their plane strides stay those of the full matrices).
int contract(const ContractArgs &c, cudaStream_t st) {
    const unsigned N = c.N;
    const size_t chunkC = c.m_pad * c.ncols; // elements of one product of this chunk
    if (c.backend == FP8) {
        // FP8: 3 piece products per modulus (9 for complex: x the three 3M products, gemmul8_real.hpp:159-180, gemmul8_complex.hpp:163-190).
        // Every product runs as its own unit on full 256 x 256 tiles (one plane pair per unit: the INT8 kernel's operand re-use and L2
        // footprint) and leaves its residue mod p as int16 in scratch; one combine pass per batch recombines them into C_mid.
        // (A 3-accumulators-per-tile kernel, EPI_F8_MOD, avoids the scratch round trip but is limited to 128-row tiles and a 3x larger
        // L2 working set: measured 1.85 vs 2.95 PFLOP/s for cuBLASLt -- see DESIGN.md.)
        // Scratch (the reference's C_hi area) holds prods x int16 x m_pad x ncols per modulus: moduli are processed in batches that fit.
        const int prods      = c.cplx ? 9 : 3;
        const size_t per_mod = (size_t)prods * sizeof(int16_t) * chunkC;
        const unsigned batch = (unsigned)std::min<size_t>(N, c.scratch_avail / per_mod);
        if (batch == 0) return G8_STATUS_NOT_SUPPORTED;
        int16_t *prod = reinterpret_cast<int16_t *>(c.scratch);
        for (unsigned u0 = 0; u0 < N; u0 += batch) {
            const unsigned nu = std::min(batch, N - u0);
            GemmArgs g{};
            g.A = c.A_lo, g.B = c.B_lo, g.strideA = c.sizeA, g.strideB = c.sizeB;
            g.m = c.m, g.n = c.ncols, g.m_pad = c.m_pad, g.k_pad = c.k_pad;
            g.num_units = (int)nu * prods, g.first_modulus = (int)u0, g.epi = EPI_F8_PROD;
            g.prods = prods, g.set_stride = (int)c.set_planes;
            g.out = prod, g.out_stride = chunkC, g.ldc = c.m_pad;
            if (int e = launch_gemm_tc(g, st)) return e;
            launch_f8_combine(prod, c.cplx, chunkC, (int)nu, (int)u0,
                              reinterpret_cast<int16_t *>(c.C_mid) + (c.cplx ? 2 : 1) * (size_t)u0 * c.mid_plane_stride, c.mid_plane_stride, st);
        }
    } else if (c.cplx) {
        // complex INT8: the three 3M products ArBr, AiBi, (Ar+Ai)(Br+Bi) of every modulus run as separate units on full 256 x 256 tiles
        // (one plane pair per unit, like the real kernel); their symmetric residues go to scratch as int8 and one combine pass per batch
        // forms {Re, Im} mod p.  The 3-accumulators-per-tile epilogue (EPI_MOD_I8_CPLX: 128-row tiles, three plane pairs live in L2 at
        // once) measured ~2.0 POP/s against ~3.2 for this path.
        const size_t per_mod = 3 * chunkC;
        const unsigned batch = (unsigned)std::min<size_t>(N, c.scratch_avail / per_mod);
        if (batch == 0) return G8_STATUS_NOT_SUPPORTED;
        int8_t *prod = c.scratch;
        for (unsigned u0 = 0; u0 < N; u0 += batch) {
            const unsigned nu = std::min(batch, N - u0);
            GemmArgs g{};
            g.A = c.A_lo, g.B = c.B_lo, g.strideA = c.sizeA, g.strideB = c.sizeB;
            g.m = c.m, g.n = c.ncols, g.m_pad = c.m_pad, g.k_pad = c.k_pad;
            g.num_units = (int)nu * 3, g.first_modulus = (int)u0, g.epi = EPI_MOD_I8, g.prods = 3;
            for (int i = 0; i < 3; ++i) g.groupA[i] = i * (int)c.set_planes + (int)u0, g.groupB[i] = i * (int)c.set_planes + (int)u0;
            g.out = prod, g.out_stride = chunkC, g.ldc = c.m_pad;
            if (int e = launch_gemm_tc(g, st)) return e;
            launch_i8_cplx_combine(prod, chunkC, (int)nu, (int)u0, c.C_mid + 2 * (size_t)u0 * c.mid_plane_stride, c.mid_plane_stride, st);
        }
    } else {
        GemmArgs g{};
        g.A = c.A_lo, g.B = c.B_lo, g.strideA = c.sizeA, g.strideB = c.sizeB;
        g.m = c.m, g.n = c.ncols, g.m_pad = c.m_pad, g.k_pad = c.k_pad;
        g.num_units = (int)N, g.first_modulus = 0, g.epi = EPI_MOD_I8;
        g.out = c.C_mid, g.out_stride = c.mid_plane_stride, g.ldc = c.m_pad;
        if (int e = launch_gemm_tc(g, st)) return e;
    }
    return 0;
}

// accurate mode: the bound GEMM over `ncols` columns with the row / column maxima fused into its epilogue (scaling_accu_real.hpp:415-432
// + :142-226; complex :444-449; FP8 find_max.hpp:82-140).  rowmax accumulates (atomicMax), colmax belongs to these columns.
int bound_gemm(bool cplx, int backend, const int8_t *A_bound, size_t sizeA, const int8_t *B_bound, size_t sizeB, size_t m, size_t ncols,
               size_t m_pad, size_t k_pad, size_t k_true, int32_t *rowmax, int32_t *colmax, cudaStream_t st) {
    GemmArgs g{};
    g.A = A_bound, g.B = B_bound, g.strideA = sizeA, g.strideB = sizeB;
    g.m = m, g.n = ncols, g.m_pad = m_pad, g.k_pad = k_pad;
    g.num_units = 1, g.first_modulus = 0;
    g.epi = backend == FP8 ? (cplx ? EPI_F8_BOUND_CPLX : EPI_F8_BOUND) : (cplx ? EPI_BOUND_MAX_CPLX : EPI_BOUND_MAX);
    g.k_true = (int)k_true;
    g.groupA[0] = 0, g.groupA[1] = 1, g.groupA[2] = 2;
    g.groupB[0] = 0, g.groupB[1] = 1, g.groupB[2] = 2;
    if (!cplx) g.groupA[1] = g.groupA[2] = g.groupB[1] = g.groupB[2] = 0;
    g.ldc = m_pad, g.rowmax = rowmax, g.colmax = colmax;
    return launch_gemm_tc(g, st);
}

// ---- helper stream for the B-side preprocessing (one per device, created on first use) ----
// fork(): the helper stream waits for everything enqueued on the caller's stream so far; join(): the caller's stream waits for the
// helper.  The fork / launch / join enqueue sequence of one call holds the device's mutex, so that concurrent g8_gemm calls of several
// host threads cannot interleave their records of the shared events.  The pattern (event-forked side stream joined back before the
// call returns) is CUDA-graph capturable.  G8_OVERLAP_SIDES=0 keeps everything on the caller's stream.
struct SideShared {
    std::mutex mu;
    cudaStream_t s   = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool ok          = false, tried = false;
};
static SideShared &side_shared() {
    static SideShared per_dev[64];
    int dev = 0;
    cudaGetDevice(&dev);
    return per_dev[dev & 63];
}
class SideStream {
    SideShared *sh_ = nullptr;
    cudaStream_t main_, use_;
    bool forked_ = false;

  public:
    SideStream(cudaStream_t main, bool want) : main_(main), use_(main) {
        static const bool enabled = [] { const char *e = getenv("G8_OVERLAP_SIDES"); return !(e && e[0] == '0'); }();
        if (!want || !enabled) return;
        SideShared &sh = side_shared();
        sh.mu.lock();
        if (!sh.tried) {
            sh.tried = true;
            sh.ok = cudaStreamCreateWithFlags(&sh.s, cudaStreamNonBlocking) == cudaSuccess &&
                    cudaEventCreateWithFlags(&sh.fork, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&sh.join, cudaEventDisableTiming) == cudaSuccess;
            if (!sh.ok) cudaGetLastError();
        }
        if (!sh.ok) {
            sh.mu.unlock();
            return;
        }
        sh_ = &sh, use_ = sh.s;
        fork();
    }
    cudaStream_t stream() const { return use_; }
    void fork() {
        if (!sh_ || forked_) return;
        cudaEventRecord(sh_->fork, main_);
        cudaStreamWaitEvent(sh_->s, sh_->fork, 0);
        forked_ = true;
    }
    void join() {
        if (!sh_ || !forked_) return;
        cudaEventRecord(sh_->join, sh_->s);
        cudaStreamWaitEvent(main_, sh_->join, 0);
        forked_ = false;
    }
    ~SideStream() {
        if (!sh_) return;
        join(); // also on the early error returns: the helper stream never outlives the call un-joined
        sh_->mu.unlock();
    }
    SideStream(const SideStream &) = delete;
    SideStream &operator=(const SideStream &) = delete;
};

SplitArgs make_split_args(int is_A, int op, size_t rows, size_t k, const void *X, size_t ld, unsigned N, int16_t *sft, int8_t *base,
                          size_t plane_stride, size_t group_stride_planes, int backend) {
    return split_args(is_A, op, rows, k, X, ld, N, sft, base, plane_stride, group_stride_planes, backend);
}
bool device_supported_cached() { return device_ok(); }

static int gemm_impl(const g8_gemm_desc &d, double *phase_ns) {
    if (!d.A || !d.B || !d.C || !d.alpha || !d.beta || !d.work) return G8_STATUS_INVALID_VALUE;
    if (d.dtype < F32 || d.dtype > C64 || d.op_A < 0 || d.op_A > 2 || d.op_B < 0 || d.op_B > 2) return G8_STATUS_INVALID_VALUE;
    if (d.num_moduli < 2 || d.num_moduli > G8_MAX_MODULI) return G8_STATUS_INVALID_VALUE;
    if (d.backend != INT8 && d.backend != FP8) return G8_STATUS_INVALID_VALUE;
    if (d.k > (size_t(1) << 17)) return G8_STATUS_INVALID_VALUE; // INT32 accumulation is exact only up to k = 2^17 (include/gemmul8.hpp:29)
    // FP8: the piece products reach 16 * 16 = 256 and are summed in binary32: exact only while 256 k <= 2^24 (SURVEY appendix B;
    // the reference documents 2^17 for both backends and silently loses exactness beyond 2^16 -- we refuse instead)
    if (d.backend == FP8 && d.k > (size_t(1) << 16)) return G8_STATUS_INVALID_VALUE;
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (d.m == 0 || d.n == 0) return 0;

    const bool cplx   = d.dtype >= C32;
    const unsigned N  = d.num_moduli;
    const bool enA = d.enable_skip_scalA, enB = d.enable_skip_scalB;
    const bool skipA = d.skip_scalA && enA, skipB = d.skip_scalB && enB;
    const Sizes s  = sizes(cplx, d.backend, d.m, d.n, d.k, N, enA, enB);
    cudaStream_t st = static_cast<cudaStream_t>(d.stream);

    // ---- workspace (gemmul8_real.hpp:95-107 / gemmul8_complex.hpp:95-118) ----
    const size_t groupA_planes = s.num_mat, groupB_planes = s.num_mat; // planes between Re / Im / Re+Im groups
    int8_t *wk  = static_cast<int8_t *>(align256(d.work));
    int8_t *wkA = d.workA ? static_cast<int8_t *>(align256(d.workA)) : nullptr;
    int8_t *wkB = d.workB ? static_cast<int8_t *>(align256(d.workB)) : nullptr;
    int8_t *A_lo  = wkA ? wkA : wk;
    int16_t *sftA = reinterpret_cast<int16_t *>(A_lo + s.sizeA * (s.num_mat + (enA ? 1 : 0)) * s.cplx3);
    int8_t *afterA = reinterpret_cast<int8_t *>(sftA + s.m_pad);
    int8_t *B_lo  = wkB ? wkB : (wkA ? wk : afterA);
    int16_t *sftB = reinterpret_cast<int16_t *>(B_lo + s.sizeB * (s.num_mat + (enB ? 1 : 0)) * s.cplx3);
    int8_t *afterB = reinterpret_cast<int8_t *>(sftB + s.n_pad);
    int8_t *C_mid = wkB ? (wkA ? wk : afterA) : afterB;
    int8_t *scratch = C_mid + s.mid * s.sizeC * N; // beyond the last C_mid plane: the reference's Lt workspace + C_hi
    const size_t scratch_avail = s.totalC - 255 - s.mid * s.sizeC * N;
    // bound planes of accurate mode: the extra plane(s) when skipping is enabled, else alias plane 0 (gemmul8_real.hpp:131-132)
    int8_t *A_bound = A_lo + (enA ? s.sizeA * s.num_mat * s.cplx3 : 0);
    int8_t *B_bound = B_lo + (enB ? s.sizeB * s.num_mat * s.cplx3 : 0);

    PhaseTimer tm(phase_ns != nullptr, st);
    tm.mark(0);

    // ---- stage 1: shifts + split ----
    if (d.k == 0) {
        cudaMemsetAsync(C_mid, 0, s.mid * s.sizeC * N, st);
        cudaMemsetAsync(sftA, 0, sizeof(int16_t) * s.m_pad, st);
        cudaMemsetAsync(sftB, 0, sizeof(int16_t) * s.n_pad, st);
    } else if (!(skipA && skipB)) {
        SplitArgs sa = split_args(1, d.op_A, d.m, d.k, d.A, d.lda, N, sftA, A_lo, s.sizeA, groupA_planes, d.backend);
        SplitArgs sb = split_args(0, d.op_B, d.n, d.k, d.B, d.ldb, N, sftB, B_lo, s.sizeB, groupB_planes, d.backend);
        // The A side and the B side are independent until they meet in a GEMM, and each of their kernels alone leaves HBM half idle
        // (they are latency / issue bound): the B side runs on a helper stream, forked from and joined back into `st`.
        SideStream side(st, !skipA && !skipB);
        cudaStream_t sB = side.stream();
        if (d.fastmode) {
            if (!skipA) launch_split(sa, d.dtype, 1, st);
            if (!skipB) launch_split(sb, d.dtype, 1, sB);
            side.join();
        } else {
            const size_t need = sizeof(int32_t) * (s.m_pad + s.n_pad);
            if (need > scratch_avail) return G8_STATUS_NOT_SUPPORTED;
            int32_t *rowmax = reinterpret_cast<int32_t *>(scratch), *colmax = rowmax + s.m_pad;
            SplitArgs ea = sa, eb = sb;
            for (int g = 0; g < 3; ++g) ea.planes[g] = A_bound + g * s.sizeA, eb.planes[g] = B_bound + g * s.sizeB;
            cudaMemsetAsync(rowmax, 0, need, st);
            side.fork();
            if (!skipA) launch_split(ea, d.dtype, 2, st);
            if (!skipB) launch_split(eb, d.dtype, 2, sB);
            side.join();
            if (int e = bound_gemm(cplx, d.backend, A_bound, s.sizeA, B_bound, s.sizeB, d.m, d.n, s.m_pad, s.k_pad, d.k, rowmax, colmax, st)) return e;
            side.fork();
            if (!skipA) {
                launch_finalize_accu_shift(sftA, rowmax, d.m, (int)N, st, d.backend);
                launch_split(sa, d.dtype, 0, st);
            }
            if (!skipB) {
                launch_finalize_accu_shift(sftB, colmax, d.n, (int)N, sB, d.backend);
                launch_split(sb, d.dtype, 0, sB);
            }
            side.join();
        }
    }
    tm.mark(1);

    // ---- stage 2: all moduli in one persistent tensor-core launch, mod-p fused ----
    if (d.k != 0) {
        ContractArgs ca{};
        ca.cplx = cplx, ca.backend = d.backend, ca.N = N, ca.m = d.m, ca.ncols = d.n, ca.m_pad = s.m_pad, ca.k_pad = s.k_pad;
        ca.A_lo = A_lo, ca.B_lo = B_lo, ca.sizeA = s.sizeA, ca.sizeB = s.sizeB, ca.set_planes = s.num_mat;
        ca.C_mid = C_mid, ca.mid_plane_stride = s.sizeC, ca.scratch = scratch, ca.scratch_avail = scratch_avail;
        if (int e = contract(ca, st)) return e;
    }
    tm.mark(2);

    // ---- stage 3: CRT + unscale + alpha/beta ----
    CrtArgs c{};
    c.C_mid = C_mid, c.ldmid = s.m_pad, c.plane_stride = s.sizeC, c.m = d.m, c.n = d.n, c.num_moduli = (int)N;
    c.C = d.C, c.ldc = d.ldc, c.sftA = sftA, c.sftB = sftB, c.alpha = d.alpha, c.beta = d.beta;
    c.backend = d.backend;
    if (int e = launch_crt(c, d.dtype, st)) return e;
    tm.mark(4);
    tm.finish(phase_ns);
    return (int)cudaPeekAtLastError();
}

// ---- FP8 backend: recombination of the per-product residues written by EPI_F8_PROD (mod.hpp:106-130, conv_hi2mid_complex.hpp:130-188) ----
// prod layout: [modulus in batch][product (3 real / 9 complex)][m_pad * n] int16.  8 consecutive elements per thread, 128-bit accesses.
__device__ __forceinline__ int32_t f8_recombine(int32_t c0, int32_t c1, int32_t c2, bool sq, int32_t sqrtp) {
    return sq ? sqrtp * (c0 + c1) + c2 : (c0 * 256) + ((c2 - c0 - c1) * 16) + c1;
}
template <bool CPLX>
__global__ void __launch_bounds__(256) f8_combine_kernel(const int16_t *__restrict__ prod, size_t elems_per_unit, int first_modulus, int16_t *__restrict__ C_mid,
                                                         size_t out_stride) {
    constexpr int NP = CPLX ? 9 : 3;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i * 8 >= elems_per_unit) return;
    const int u = blockIdx.y, midx = first_modulus + u;
    const int32_t p = g8d_moduli[FP8][midx], pinv = g8d_pinv32[FP8][midx];
    const bool sq = midx < 6;
    const int32_t sqrtp = sq ? g8d_f8sqrt[midx] : 0;
    const int16_t *src = prod + (size_t)u * NP * elems_per_unit + i * 8;
    uint32_t w[NP][4];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(src + (size_t)q * elems_per_unit));
        w[q][0] = v.x, w[q][1] = v.y, w[q][2] = v.z, w[q][3] = v.w;
    }
    auto get = [&](int q, int j) { return (int32_t)(int16_t)(w[q][j >> 1] >> ((j & 1) * 16)); };
    if constexpr (!CPLX) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const int32_t a = mod_i32(f8_recombine(get(0, j), get(1, j), get(2, j), sq, sqrtp), p, pinv);
            const int32_t b = mod_i32(f8_recombine(get(0, j + 1), get(1, j + 1), get(2, j + 1), sq, sqrtp), p, pinv);
            o[j >> 1]       = (uint32_t)(a & 0xFFFF) | ((uint32_t)b << 16);
        }
        *reinterpret_cast<uint4 *>(C_mid + (size_t)u * out_stride + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            // every 3M product is first reduced mod p (the recombined value is < 2^21, the differences stay far below 2^31)
            const int32_t rr = mod_i32(f8_recombine(get(0, j), get(1, j), get(2, j), sq, sqrtp), p, pinv);
            const int32_t ii = mod_i32(f8_recombine(get(3, j), get(4, j), get(5, j), sq, sqrtp), p, pinv);
            const int32_t ss = mod_i32(f8_recombine(get(6, j), get(7, j), get(8, j), sq, sqrtp), p, pinv);
            const int32_t re = mod_i32(rr - ii, p, pinv), im = mod_i32(ss - rr - ii, p, pinv);
            o[j]             = (uint32_t)(re & 0xFFFF) | ((uint32_t)im << 16);
        }
        uint4 *dst = reinterpret_cast<uint4 *>(C_mid + ((size_t)u * out_stride + i * 8) * 2);
        dst[0]     = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1]     = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// ---- INT8 backend, complex: 3M recombination of the per-product residues (conv_hi2mid_complex.hpp:46-127); 16 elements per thread ----
__global__ void __launch_bounds__(256) i8_cplx_combine_kernel(const int8_t *__restrict__ prod, size_t elems_per_unit, int first_modulus, int8_t *__restrict__ C_mid,
                                                              size_t out_stride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i * 16 >= elems_per_unit) return;
    const int u = blockIdx.y, midx = first_modulus + u;
    const int32_t p = g8d_moduli[INT8][midx], pinv = g8d_pinv32[INT8][midx];
    const int8_t *src = prod + (size_t)u * 3 * elems_per_unit + i * 16;
    const int4 a = __ldcs(reinterpret_cast<const int4 *>(src)), b = __ldcs(reinterpret_cast<const int4 *>(src + elems_per_unit)),
               c = __ldcs(reinterpret_cast<const int4 *>(src + 2 * elems_per_unit));
    const int aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, cw[4] = {c.x, c.y, c.z, c.w};
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        uint32_t word = 0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int sh = 8 * ((j + e) & 3), wi = (j + e) >> 2;
            const int32_t rr = (int32_t)(int8_t)(aw[wi] >> sh), ii = (int32_t)(int8_t)(bw[wi] >> sh), ss = (int32_t)(int8_t)(cw[wi] >> sh);
            const int32_t re = mod_i32(rr - ii, p, pinv), im = mod_i32(ss - rr - ii, p, pinv);
            word |= ((uint32_t)(re & 0xFF) | ((uint32_t)(im & 0xFF) << 8)) << (16 * e);
        }
        o[j >> 1] = word;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(C_mid + ((size_t)u * out_stride + i * 16) * 2);
    dst[0]     = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1]     = make_uint4(o[4], o[5], o[6], o[7]);
}

// K-sharded complex INT8: the three 3M products of every modulus arrive as `nparts` per-shard residue arrays
// [shard][modulus][rr, ii, ss][elems]; sum over the shards (dp4a against one-hot selectors), reduce mod p, then the 3M recombination
__global__ void __launch_bounds__(256) i8_cplx_combine_parts_kernel(const int8_t *__restrict__ parts, int nparts, size_t part_stride, size_t elems_per_unit,
                                                                    int first_modulus, int8_t *__restrict__ C_mid, size_t out_stride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i * 16 >= elems_per_unit) return;
    const int u = blockIdx.y, midx = first_modulus + u;
    const int32_t p = g8d_moduli[INT8][midx], pinv = g8d_pinv32[INT8][midx];
    const int8_t *src = parts + (size_t)u * 3 * elems_per_unit + i * 16;
    int32_t acc[3][16];
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[q][j] = 0;
    for (int s = 0; s < nparts; ++s) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int4 v   = __ldcs(reinterpret_cast<const int4 *>(src + (size_t)s * part_stride + (size_t)q * elems_per_unit));
            const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[q][j] = __dp4a(w[j >> 2], 1 << (8 * (j & 3)), acc[q][j]);
        }
    }
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        uint32_t word = 0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int32_t rr = mod_i32(acc[0][j + e], p, pinv), ii = mod_i32(acc[1][j + e], p, pinv), ss = mod_i32(acc[2][j + e], p, pinv);
            const int32_t re = mod_i32(rr - ii, p, pinv), im = mod_i32(ss - rr - ii, p, pinv);
            word |= ((uint32_t)(re & 0xFF) | ((uint32_t)(im & 0xFF) << 8)) << (16 * e);
        }
        o[j >> 1] = word;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(C_mid + ((size_t)u * out_stride + i * 16) * 2);
    dst[0]     = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1]     = make_uint4(o[4], o[5], o[6], o[7]);
}
void launch_i8_cplx_combine_parts(const int8_t *parts, int nparts, size_t part_stride, size_t elems_per_unit, int num_units, int first_modulus, int8_t *C_mid,
                                  size_t out_stride, cudaStream_t st) {
    const size_t groups = elems_per_unit / 16;
    const dim3 grid((unsigned)((groups + 255) / 256), (unsigned)num_units);
    i8_cplx_combine_parts_kernel<<<grid, 256, 0, st>>>(parts, nparts, part_stride, elems_per_unit, first_modulus, C_mid, out_stride);
}

// K-sharded FP8 (real and complex): sum of the per-shard C_mid arrays (each already a canonical residue of that shard's partial product),
// reduced again: mod_i32 maps every member of a residue class to the same representative in (-p/2, p/2], so the owner's C_mid is
// bit-identical to what the single-GPU call computes from the un-sharded products.  8 int16 per thread, 128-bit accesses.
__global__ void __launch_bounds__(256) i16_sum_parts_kernel(const int16_t *__restrict__ parts, int nparts, size_t part_stride, size_t elems_per_unit,
                                                            int first_modulus, int16_t *__restrict__ C_mid, size_t out_stride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i * 8 >= elems_per_unit) return;
    const int u = blockIdx.y, midx = first_modulus + u;
    const int32_t p = g8d_moduli[FP8][midx], pinv = g8d_pinv32[FP8][midx];
    const int16_t *src = parts + (size_t)u * elems_per_unit + i * 8;
    int32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = 0; s < nparts; ++s) {
        const uint4 v       = __ldcs(reinterpret_cast<const uint4 *>(src + (size_t)s * part_stride));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += (int32_t)(int16_t)(w[j >> 1] >> ((j & 1) * 16));
    }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const int32_t a = mod_i32(acc[j], p, pinv), b = mod_i32(acc[j + 1], p, pinv);
        o[j >> 1]       = (uint32_t)(a & 0xFFFF) | ((uint32_t)b << 16);
    }
    *reinterpret_cast<uint4 *>(C_mid + (size_t)u * out_stride + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}
void launch_i16_sum_parts(const int16_t *parts, int nparts, size_t part_stride, size_t elems_per_unit, int num_units, int first_modulus, int16_t *C_mid,
                          size_t out_stride, cudaStream_t st) {
    if (elems_per_unit == 0 || num_units == 0) return;
    const dim3 grid((unsigned)((elems_per_unit / 8 + 255) / 256), (unsigned)num_units);
    i16_sum_parts_kernel<<<grid, 256, 0, st>>>(parts, nparts, part_stride, elems_per_unit, first_modulus, C_mid, out_stride);
}

void launch_i8_cplx_combine(const int8_t *prod, size_t elems_per_unit, int num_units, int first_modulus, int8_t *C_mid, size_t out_stride,
                            cudaStream_t st) {
    const size_t groups = elems_per_unit / 16;
    const dim3 grid((unsigned)((groups + 255) / 256), (unsigned)num_units);
    i8_cplx_combine_kernel<<<grid, 256, 0, st>>>(prod, elems_per_unit, first_modulus, C_mid, out_stride);
}

void launch_f8_combine(const int16_t *prod, bool cplx, size_t elems_per_unit, int num_units, int first_modulus, int16_t *C_mid, size_t out_stride,
                       cudaStream_t st) {
    const size_t groups = elems_per_unit / 8; // m_pad * ncols, m_pad % 256 == 0
    const dim3 grid((unsigned)((groups + 255) / 256), (unsigned)num_units);
    if (cplx) f8_combine_kernel<true><<<grid, 256, 0, st>>>(prod, elems_per_unit, first_modulus, C_mid, out_stride);
    else f8_combine_kernel<false><<<grid, 256, 0, st>>>(prod, elems_per_unit, first_modulus, C_mid, out_stride);
}

// ---- K-sharded multi-GPU helpers (no reference counterpart; arithmetic = conv_hi2mid_real.hpp:19-22) ----
// C_hi (int32, summed over K-shards) -> symmetric residues.  4 consecutive rows per thread.
__global__ void requant_i32_kernel(const int32_t *__restrict__ C_hi, size_t rows4, size_t cols, size_t in_ld, size_t in_unit_stride,
                                   int first_modulus, int8_t *__restrict__ C_mid, size_t out_ld, size_t out_unit_stride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows4 * cols) return;
    const size_t col = i / rows4, r4 = (i - col * rows4) * 4;
    const int u     = blockIdx.y;
    const int32_t p = g8d_moduli[INT8][first_modulus + u], pinv = g8d_pinv32[INT8][first_modulus + u];
    const int4 v    = *reinterpret_cast<const int4 *>(C_hi + (size_t)u * in_unit_stride + col * in_ld + r4);
    const int32_t a = mod_i32(v.x, p, pinv), b = mod_i32(v.y, p, pinv), c = mod_i32(v.z, p, pinv), e = mod_i32(v.w, p, pinv);
    *reinterpret_cast<uint32_t *>(C_mid + (size_t)u * out_unit_stride + col * out_ld + r4) =
        (uint32_t)(a & 0xFF) | ((uint32_t)(b & 0xFF) << 8) | ((uint32_t)(c & 0xFF) << 16) | ((uint32_t)e << 24);
}
// sum of `nparts` int8 residue arrays (one per K-shard, identical layout) reduced mod p again.  16 consecutive rows per thread:
// one 128-bit load per part, bytes accumulated with dp4a against one-hot selectors (sign-extending add of byte j in ONE instruction).
__global__ void __launch_bounds__(256) residue_sum_kernel(const int8_t *__restrict__ parts, int nparts, size_t part_stride, size_t rows16, size_t cols,
                                                          size_t in_ld, size_t in_unit_stride, int first_modulus, int8_t *__restrict__ C_mid, size_t out_ld,
                                                          size_t out_unit_stride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows16 * cols) return;
    const size_t col = i / rows16, r0 = (i - col * rows16) * 16;
    const int u     = blockIdx.y;
    const int32_t p = g8d_moduli[INT8][first_modulus + u], pinv = g8d_pinv32[INT8][first_modulus + u];
    int32_t acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0;
    const int8_t *src = parts + (size_t)u * in_unit_stride + col * in_ld + r0;
    for (int q = 0; q < nparts; ++q) {
        const int4 v = __ldcs(reinterpret_cast<const int4 *>(src + (size_t)q * part_stride));
        const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = __dp4a(w[j >> 2], 1 << (8 * (j & 3)), acc[j]);
    }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int32_t a = mod_i32(acc[4 * j], p, pinv), b = mod_i32(acc[4 * j + 1], p, pinv), c = mod_i32(acc[4 * j + 2], p, pinv),
                      e = mod_i32(acc[4 * j + 3], p, pinv);
        o[j] = (uint32_t)(a & 0xFF) | ((uint32_t)(b & 0xFF) << 8) | ((uint32_t)(c & 0xFF) << 16) | ((uint32_t)e << 24);
    }
    *reinterpret_cast<uint4 *>(C_mid + (size_t)u * out_unit_stride + col * out_ld + r0) = make_uint4(o[0], o[1], o[2], o[3]);
}
// row / column maxima of an int32 slab (the reduced bound product of accurate mode); one block per column
__global__ void maxabs_i32_kernel(const int32_t *__restrict__ C, size_t rows, size_t ld, int32_t *__restrict__ rowmax, int32_t *__restrict__ colmax) {
    const int32_t *col = C + (size_t)blockIdx.x * ld;
    int32_t cm = 0;
    for (size_t r = threadIdx.x; r < rows; r += blockDim.x) {
        const int32_t v = col[r];
        cm = max(cm, v);
        if (v > 0) atomicMax(&rowmax[r], v);
    }
    cm = __reduce_max_sync(0xffffffffu, cm);
    if ((threadIdx.x & 31) == 0 && cm > 0) atomicMax(&colmax[blockIdx.x], cm);
}

// bound product of accurate mode, K-sharded: sum of `nparts` int32 slabs (one per K-shard, written by the fused GEMM -> scatter),
// then row / column maxima of the sum
__global__ void __launch_bounds__(256) maxabs_i32_parts_kernel(const int32_t *__restrict__ parts, int nparts, size_t part_stride, size_t rows, size_t ld,
                                                               int32_t *__restrict__ rowmax, int32_t *__restrict__ colmax) {
    const int32_t *col = parts + (size_t)blockIdx.x * ld;
    int32_t cm = 0;
    for (size_t r = (size_t)threadIdx.x * 4; r < rows; r += (size_t)blockDim.x * 4) { // ld and rows' padding are multiples of 4 (m_pad)
        int4 v = make_int4(0, 0, 0, 0);
        for (int q = 0; q < nparts; ++q) {
            const int4 w = __ldcs(reinterpret_cast<const int4 *>(col + (size_t)q * part_stride + r));
            v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;
        }
        const int32_t e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (r + j < rows) {
                cm = max(cm, e[j]);
                if (e[j] > 0 && e[j] > *reinterpret_cast<volatile int32_t *>(&rowmax[r + j])) atomicMax(&rowmax[r + j], e[j]); // plain pre-check saves most atomics
            }
    }
    cm = __reduce_max_sync(0xffffffffu, cm);
    if ((threadIdx.x & 31) == 0 && cm > 0) atomicMax(&colmax[blockIdx.x], cm);
}

} // namespace g8

using namespace g8;

extern "C" {

__attribute__((visibility("default"))) size_t g8_work_size(int is_complex, int backend, size_t m, size_t n, size_t k, unsigned num_moduli, int enA, int enB,
                    size_t *workSizeA, size_t *workSizeB) {
    const Sizes s = sizes(is_complex != 0, backend, m, n, k, num_moduli, enA != 0, enB != 0);
    if (workSizeA) *workSizeA = s.totalA;
    if (workSizeB) *workSizeB = s.totalB;
    return s.totalA + s.totalB + s.totalC;
}

__attribute__((visibility("default"))) int g8_gemm(const g8_gemm_desc *d, double *phase_ns) {
    if (!d) return G8_STATUS_INVALID_VALUE;
    if (phase_ns) phase_ns[0] = phase_ns[1] = phase_ns[2] = phase_ns[3] = 0.0;
    return gemm_impl(*d, phase_ns);
}

__attribute__((visibility("default"))) int g8_stage_split(int dtype, int is_A, int op, size_t rows, size_t k, const void *X, size_t ld, unsigned num_moduli, int mode,
                   int16_t *sft, int8_t *planes, size_t plane_stride_bytes, size_t group_stride_planes, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!X || !sft || !planes || dtype < F32 || dtype > C64 || num_moduli < 2 || num_moduli > G8_MAX_MODULI) return G8_STATUS_INVALID_VALUE;
    if (rows == 0 || k == 0) return 0;
    if (k > (size_t(1) << 17)) return G8_STATUS_INVALID_VALUE; // same bound as g8_gemm (include/gemmul8.hpp:29)
    SplitArgs a = split_args(is_A, op, rows, k, X, ld, num_moduli, sft, planes, plane_stride_bytes, group_stride_planes);
    if (mode >= 2)
        for (int g = 0; g < 3; ++g) a.planes[g] = planes + g * plane_stride_bytes;
    launch_split(a, dtype, mode, static_cast<cudaStream_t>(stream));
    return (int)cudaGetLastError();
}

__attribute__((visibility("default"))) int g8_stage_finalize_shift(int16_t *sft, const int32_t *cmax, size_t count, unsigned num_moduli, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (count == 0) return 0;
    launch_finalize_accu_shift(sft, cmax, count, (int)num_moduli, static_cast<cudaStream_t>(stream));
    return (int)cudaGetLastError();
}

__attribute__((visibility("default"))) int g8_stage_gemm(int epilogue, int use_simt, const int8_t *A_lo, size_t strideA, const int8_t *B_lo, size_t strideB, size_t m, size_t n,
                  size_t k_pad, int num_units, int first_modulus, const int *groupA, const int *groupB, void *out, size_t out_stride,
                  size_t ldc, int32_t *rowmax, int32_t *colmax, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!A_lo || !B_lo || k_pad % 256 || epilogue < 0 || epilogue > 7) return G8_STATUS_INVALID_VALUE;
    if (k_pad > (size_t(1) << 17)) return G8_STATUS_INVALID_VALUE; // |sum| <= k * 2^14 must fit INT32 (and f32 exactly for the FP8 pieces: callers keep k <= 2^16)
    GemmArgs g{};
    g.A = A_lo, g.B = B_lo, g.strideA = strideA, g.strideB = strideB, g.m = m, g.n = n, g.m_pad = pad256(m), g.k_pad = k_pad;
    g.num_units = num_units, g.first_modulus = first_modulus, g.epi = epilogue;
    for (int i = 0; i < 3; ++i) g.groupA[i] = groupA ? groupA[i] : 0, g.groupB[i] = groupB ? groupB[i] : 0;
    g.out = out, g.out_stride = out_stride, g.ldc = ldc, g.rowmax = rowmax, g.colmax = colmax;
    g.k_true = (int)k_pad;
    if (use_simt && epilogue > 4) return G8_STATUS_NOT_SUPPORTED;
    return use_simt ? launch_gemm_simt(g, static_cast<cudaStream_t>(stream)) : launch_gemm_tc(g, static_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int g8_stage_crt(int dtype, const void *C_mid, size_t ldmid, size_t plane_stride, size_t m, size_t n, unsigned num_moduli, void *C,
                 size_t ldc, const int16_t *sftA, const int16_t *sftB, const void *alpha, const void *beta, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!C_mid || !C || !sftA || !sftB || !alpha || !beta || num_moduli < 2 || num_moduli > G8_MAX_MODULI) return G8_STATUS_INVALID_VALUE;
    if (ldmid % 8 || plane_stride % 8 || reinterpret_cast<uintptr_t>(C_mid) % 8) return G8_STATUS_INVALID_VALUE; // 64-bit residue loads
    CrtArgs c{};
    c.C_mid = C_mid, c.ldmid = ldmid, c.plane_stride = plane_stride, c.m = m, c.n = n, c.num_moduli = (int)num_moduli;
    c.C = C, c.ldc = ldc, c.sftA = sftA, c.sftB = sftB, c.alpha = alpha, c.beta = beta;
    return launch_crt(c, dtype, static_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int g8_stage_crt_parts(int dtype, const void *parts, int nparts, size_t part_stride, size_t ldmid, size_t plane_stride, size_t m, size_t n,
                                                               unsigned num_moduli, void *C, size_t ldc, const int16_t *sftA, const int16_t *sftB, const void *alpha,
                                                               const void *beta, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!parts || !C || !sftA || !sftB || !alpha || !beta || num_moduli < 2 || num_moduli > G8_MAX_MODULI) return G8_STATUS_INVALID_VALUE;
    if (nparts < 1 || nparts > G8_MAX_PEERS || part_stride % 8 || (dtype != F32 && dtype != F64)) return G8_STATUS_INVALID_VALUE;
    if (ldmid % 8 || plane_stride % 8 || reinterpret_cast<uintptr_t>(parts) % 8) return G8_STATUS_INVALID_VALUE; // 64-bit residue loads
    CrtArgs c{};
    c.C_mid = parts, c.ldmid = ldmid, c.plane_stride = plane_stride, c.m = m, c.n = n, c.num_moduli = (int)num_moduli;
    c.C = C, c.ldc = ldc, c.sftA = sftA, c.sftB = sftB, c.alpha = alpha, c.beta = beta;
    c.nparts = nparts, c.part_stride = part_stride, c.backend = INT8;
    return launch_crt(c, dtype, static_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int g8_stage_requant_i32(const int32_t *C_hi, size_t rows, size_t cols, size_t in_ld, size_t in_unit_stride, int num_units,
                         int first_modulus, int8_t *C_mid, size_t out_ld, size_t out_unit_stride, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!C_hi || !C_mid || rows % 4 || in_ld % 4 || out_ld % 4 || in_unit_stride % 4 || out_unit_stride % 4) return G8_STATUS_INVALID_VALUE;
    if (rows == 0 || cols == 0 || num_units == 0) return 0;
    const size_t r4 = rows / 4;
    const dim3 grid((unsigned)((r4 * cols + 255) / 256), (unsigned)num_units);
    requant_i32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(C_hi, r4, cols, in_ld, in_unit_stride, first_modulus, C_mid, out_ld,
                                                                            out_unit_stride);
    return (int)cudaGetLastError();
}

__attribute__((visibility("default"))) int g8_stage_residue_sum(const int8_t *parts, int nparts, size_t part_stride, size_t rows, size_t cols, size_t in_ld, size_t in_unit_stride,
                         int num_units, int first_modulus, int8_t *C_mid, size_t out_ld, size_t out_unit_stride, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!parts || !C_mid || nparts < 1 || nparts > 64 || rows % 16 || in_ld % 16 || out_ld % 16 || in_unit_stride % 16 || out_unit_stride % 16 ||
        part_stride % 16 || reinterpret_cast<uintptr_t>(parts) % 16 || reinterpret_cast<uintptr_t>(C_mid) % 16)
        return G8_STATUS_INVALID_VALUE;
    if (rows == 0 || cols == 0 || num_units == 0) return 0;
    const size_t r16 = rows / 16;
    const dim3 grid((unsigned)((r16 * cols + 255) / 256), (unsigned)num_units);
    residue_sum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(parts, nparts, part_stride, r16, cols, in_ld, in_unit_stride, first_modulus,
                                                                            C_mid, out_ld, out_unit_stride);
    return (int)cudaGetLastError();
}

__attribute__((visibility("default"))) int g8_stage_maxabs_i32(const int32_t *C, size_t rows, size_t cols, size_t ld, int32_t *rowmax, int32_t *colmax, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!C || !rowmax || !colmax) return G8_STATUS_INVALID_VALUE;
    if (rows == 0 || cols == 0) return 0;
    maxabs_i32_kernel<<<(unsigned)cols, 256, 0, static_cast<cudaStream_t>(stream)>>>(C, rows, ld, rowmax, colmax);
    return (int)cudaGetLastError();
}

__attribute__((visibility("default"))) int g8_stage_maxabs_i32_parts(const int32_t *parts, int nparts, size_t part_stride, size_t rows, size_t cols, size_t ld,
                                                                      int32_t *rowmax, int32_t *colmax, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!parts || !rowmax || !colmax || nparts < 1 || nparts > G8_MAX_PEERS || ld % 4 || part_stride % 4 || reinterpret_cast<uintptr_t>(parts) % 16)
        return G8_STATUS_INVALID_VALUE;
    if (rows == 0 || cols == 0) return 0;
    maxabs_i32_parts_kernel<<<(unsigned)cols, 256, 0, static_cast<cudaStream_t>(stream)>>>(parts, nparts, part_stride, rows, ld, rowmax, colmax);
    return (int)cudaGetLastError();
}

// bound GEMM of accurate mode over `chain` K-slabs: plane c of A_planes / B_planes (strideA / strideB bytes apart) holds K-slab c of the
// int8 bound matrices; the accumulator sums all slabs, the epilogue reduces row / column maxima (never written).  INT8 backend, real.
__attribute__((visibility("default"))) int g8_stage_gemm_bound_chain(const int8_t *A_planes, size_t strideA, const int8_t *B_planes, size_t strideB, size_t m, size_t n,
                                                                      size_t k_pad, int chain, int32_t *rowmax, int32_t *colmax, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!A_planes || !B_planes || !rowmax || !colmax || k_pad % 256 || chain < 1 || chain > 64) return G8_STATUS_INVALID_VALUE;
    if ((size_t)chain * k_pad > (size_t(1) << 19)) return G8_STATUS_INVALID_VALUE; // bound entries <= 64: |sum| <= K * 2^12 < 2^31
    GemmArgs g{};
    g.A = A_planes, g.B = B_planes, g.strideA = strideA, g.strideB = strideB, g.m = m, g.n = n, g.m_pad = pad256(m), g.k_pad = k_pad;
    g.num_units = 1, g.first_modulus = 0, g.epi = EPI_BOUND_MAX, g.kchain = chain;
    g.ldc = pad256(m), g.rowmax = rowmax, g.colmax = colmax, g.k_true = (int)k_pad;
    return launch_gemm_tc(g, static_cast<cudaStream_t>(stream));
}

// ---- fused GEMM -> scatter over peer memory (K-sharded multi-GPU) ----
__attribute__((visibility("default"))) int g8_stage_gemm_scatter(int epilogue, const int8_t *A_lo, size_t strideA, const int8_t *B_lo, size_t strideB, size_t m, size_t n,
                                                                  size_t k_pad, int num_units, int first_modulus, void *const *peer_out, int world, int rank,
                                                                  size_t out_stride, size_t ldc, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!A_lo || !B_lo || !peer_out || k_pad % 256 || k_pad > (size_t(1) << 17) || (epilogue != EPI_MOD_I8 && epilogue != EPI_RAW_I32)) return G8_STATUS_INVALID_VALUE;
    if (world < 1 || world > G8_MAX_PEERS || rank < 0 || rank >= world || n % (size_t)world || (n / (size_t)world) % 256) return G8_STATUS_INVALID_VALUE;
    GemmArgs g{};
    g.A = A_lo, g.B = B_lo, g.strideA = strideA, g.strideB = strideB, g.m = m, g.n = n, g.m_pad = pad256(m), g.k_pad = k_pad;
    g.num_units = num_units, g.first_modulus = first_modulus, g.epi = epilogue;
    g.out = nullptr, g.out_stride = out_stride, g.ldc = ldc, g.k_true = (int)k_pad;
    for (int i = 0; i < world; ++i) {
        if (!peer_out[i]) return G8_STATUS_INVALID_VALUE;
        g.peer_out[i] = peer_out[i];
    }
    g.owner_cols = n / (size_t)world, g.rank = rank, g.world = world;
    return launch_gemm_tc(g, static_cast<cudaStream_t>(stream));
}

// ---- device buffers that can be mapped into the other ranks of the node (CUDA IPC) ----
__attribute__((visibility("default"))) int g8_peer_alloc(size_t bytes, void **dptr, void *handle64) {
    if (!dptr || !handle64 || bytes == 0) return G8_STATUS_INVALID_VALUE;
    cudaError_t e = cudaMalloc(dptr, bytes);
    if (e != cudaSuccess) return (int)e;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle64), *dptr);
    if (e != cudaSuccess) cudaFree(*dptr), *dptr = nullptr;
    return (int)e;
}
__attribute__((visibility("default"))) int g8_peer_open(const void *handle64, void **dptr) {
    if (!dptr || !handle64) return G8_STATUS_INVALID_VALUE;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    return (int)cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess);
}
__attribute__((visibility("default"))) int g8_peer_close(void *dptr) { return dptr ? (int)cudaIpcCloseMemHandle(dptr) : 0; }
__attribute__((visibility("default"))) int g8_peer_free(void *dptr) { return dptr ? (int)cudaFree(dptr) : 0; }

__attribute__((visibility("default"))) int g8_stage_stats(int dtype, int is_A, int op, size_t rows, size_t k, const void *X, size_t ld, double *amax, double *sumsq, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!X || !amax || !sumsq || dtype < F32 || dtype > C64) return G8_STATUS_INVALID_VALUE;
    if (rows == 0) return 0;
    SplitArgs a = split_args(is_A, op, rows, k, X, ld, 2, nullptr, nullptr, 0, 0);
    launch_stats(a, dtype, amax, sumsq, static_cast<cudaStream_t>(stream));
    return (int)cudaGetLastError();
}

__attribute__((visibility("default"))) int g8_stage_shift_from_stats(const double *amax, const double *sumsq, size_t count, unsigned num_moduli, int kind, int16_t *sft, void *stream) {
    if (!device_ok()) return G8_STATUS_NO_DEVICE_CODE;
    if (!amax || !sft || (kind == 0 && !sumsq) || num_moduli < 2 || num_moduli > G8_MAX_MODULI) return G8_STATUS_INVALID_VALUE;
    if (count == 0) return 0;
    launch_shift_from_stats(amax, sumsq, count, (int)num_moduli, kind, sft, static_cast<cudaStream_t>(stream));
    return (int)cudaGetLastError();
}

__attribute__((visibility("default"))) const char *g8_version(void) { return "gemmul8_b200 0.1 (sm_100a, tcgen05 kind::i8)"; }

__attribute__((visibility("default"))) int g8_device_supported(int device) {
    int major = 0, minor = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    return major == 10 && minor == 0;
}

} // extern "C"
