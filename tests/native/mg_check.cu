// Native multi-GPU drop-in check: one process per GPU (fork), NO Python, NO NCCL / MPI.  Every rank drives the C ABI of
// include/gemmul8_c.h -- g8_mg_comm_create / connect, g8_mg_plan_create, g8_gemm_mg -- on its K-slab of a synthetic DGEMM / SGEMM,
// then recomputes the WHOLE product with the single-GPU g8_gemm on the concatenated operands and compares its column slab:
// accurate mode must match bit for bit, fast mode to 1e-9 (see include/gemmul8_c.h).  The 64-byte IPC handles travel through a shared
// anonymous mapping created before the fork (any transport would do).
//   usage: mg_check [world]      (default: all visible GPUs, at most 8; needs >= 2)
#include "../../include/gemmul8_c.h"

#include <cuComplex.h>
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Shared {
    std::atomic<int> arrived[4];
    unsigned char handles[8][64];
    int ok[8];
};
static void host_barrier(Shared *sh, int idx, int world) {
    sh->arrived[idx].fetch_add(1);
    while (sh->arrived[idx].load() < world) usleep(100);
}
#define CK(x)                                                                      \
    do {                                                                           \
        const int _e = (int)(x);                                                   \
        if (_e != 0) {                                                             \
            std::fprintf(stderr, "rank %d: %s -> %d (line %d)\n", rank, #x, _e, __LINE__); \
            return 1;                                                              \
        }                                                                          \
    } while (0)

template <typename T> static int run_case(g8_mg_comm *comm, int rank, int world, int dtype, unsigned N, int fast, int opA, int opB, int backend = G8_BACKEND_INT8) {
    const size_t m = 300, n = 256 * (size_t)world, kl = 384, K = kl * world, nc = n / world;
    // full operands, identical on every rank (same seeds): A stored as op_A wants it, likewise B
    const size_t rA = opA == G8_OP_N ? m : K, cA = opA == G8_OP_N ? K : m, rB = opB == G8_OP_N ? K : n, cB = opB == G8_OP_N ? n : K;
    T *A, *B, *Cfull, *Cslab, *Al, *Bl;
    CK(cudaMalloc(&A, sizeof(T) * rA * cA));
    CK(cudaMalloc(&B, sizeof(T) * rB * cB));
    CK(cudaMalloc(&Cfull, sizeof(T) * m * n));
    CK(cudaMalloc(&Cslab, sizeof(T) * m * nc));
    CK(g8_randmat(dtype, A, rA, cA, 0.5, 11, nullptr));
    CK(g8_randmat(dtype, B, rB, cB, 0.5, 22, nullptr));
    // this rank's K-slab: op(A)[:, K_r] and op(B)[K_r, :] as compact column-major matrices
    const size_t rAl = opA == G8_OP_N ? m : kl, cAl = opA == G8_OP_N ? kl : m, rBl = opB == G8_OP_N ? kl : n, cBl = opB == G8_OP_N ? n : kl;
    CK(cudaMalloc(&Al, sizeof(T) * rAl * cAl));
    CK(cudaMalloc(&Bl, sizeof(T) * rBl * cBl));
    if (opA == G8_OP_N) { CK(cudaMemcpy(Al, A + (size_t)rank * kl * m, sizeof(T) * m * kl, cudaMemcpyDeviceToDevice)); } // columns K_r
    else { CK(cudaMemcpy2D(Al, sizeof(T) * kl, A + (size_t)rank * kl, sizeof(T) * K, sizeof(T) * kl, m, cudaMemcpyDeviceToDevice)); } // rows K_r
    if (opB == G8_OP_N) { CK(cudaMemcpy2D(Bl, sizeof(T) * kl, B + (size_t)rank * kl, sizeof(T) * K, sizeof(T) * kl, n, cudaMemcpyDeviceToDevice)); } // rows K_r
    else { CK(cudaMemcpy(Bl, B + (size_t)rank * kl * n, sizeof(T) * n * kl, cudaMemcpyDeviceToDevice)); } // columns K_r
    T one, zero;
    std::memset(&one, 0, sizeof(T)), std::memset(&zero, 0, sizeof(T));
    if (dtype == G8_R32F || dtype == G8_C32F) *reinterpret_cast<float *>(&one) = 1.0f;
    else *reinterpret_cast<double *>(&one) = 1.0;
    const bool cplx = dtype >= G8_C32F;
    // single-GPU reference on the full K
    const size_t wbytes = g8_work_size(cplx, backend, m, n, K, N, 0, 0, nullptr, nullptr);
    void *work;
    CK(cudaMalloc(&work, wbytes));
    g8_gemm_desc d{};
    d.dtype = dtype, d.backend = backend, d.op_A = opA, d.op_B = opB, d.m = m, d.n = n, d.k = K;
    d.alpha = &one, d.A = A, d.lda = rA, d.B = B, d.ldb = rB, d.beta = &zero, d.C = Cfull, d.ldc = m, d.num_moduli = N, d.fastmode = fast, d.work = work;
    CK(g8_gemm(&d, nullptr));
    // sharded
    g8_mg_plan *plan = nullptr;
    if (backend == G8_BACKEND_INT8) CK(g8_mg_plan_create(&plan, comm, dtype, opA, opB, m, n, kl, N, fast));
    else CK(g8_mg_plan_create_backend(&plan, comm, dtype, backend, opA, opB, m, n, kl, N, fast));
    for (int rep = 0; rep < 2; ++rep) CK(g8_gemm_mg(plan, &one, Al, rAl, Bl, rBl, &zero, Cslab, m, nullptr)); // twice: the receive areas are re-used
    CK(cudaDeviceSynchronize());
    CK(g8_mg_comm_status(comm));
    std::vector<T> want(m * nc), got(m * nc);
    CK(cudaMemcpy(want.data(), Cfull + (size_t)rank * nc * m, sizeof(T) * m * nc, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(got.data(), Cslab, sizeof(T) * m * nc, cudaMemcpyDeviceToHost));
    const bool same = std::memcmp(want.data(), got.data(), sizeof(T) * m * nc) == 0;
    double num = 0, den = 0;
    const bool dbl = dtype == G8_R64F || dtype == G8_C64F;
    const size_t scalars = m * nc * (cplx ? 2 : 1);
    for (size_t i = 0; i < scalars; ++i) {
        const double w = dbl ? reinterpret_cast<const double *>(want.data())[i] : (double)reinterpret_cast<const float *>(want.data())[i];
        const double g = dbl ? reinterpret_cast<const double *>(got.data())[i] : (double)reinterpret_cast<const float *>(got.data())[i];
        num = std::fmax(num, std::fabs(w - g)), den = std::fmax(den, std::fabs(w));
    }
    const bool ok = fast ? (num <= (dbl ? 1e-9 : 1e-3) * den) : same;
    std::printf("rank %d: %cGEMM %s N=%u %s op%d%d  %s (max diff / max = %.2e)\n", rank, "SDCZ"[dtype], backend == G8_BACKEND_INT8 ? "INT8" : "FP8", N,
                fast ? "fast" : "accu", opA, opB,
                same ? "bit-identical" : (ok ? "within tolerance" : "MISMATCH"), den > 0 ? num / den : 0.0);
    CK(g8_mg_plan_destroy(plan));
    cudaFree(A), cudaFree(B), cudaFree(Cfull), cudaFree(Cslab), cudaFree(Al), cudaFree(Bl), cudaFree(work);
    return ok ? 0 : 1;
}

static int rank_main(Shared *sh, int rank, int world) {
    CK(cudaSetDevice(rank));
    g8_mg_comm *comm = nullptr;
    CK(g8_mg_comm_create(&comm, world, rank, 1 << 20, sh->handles[rank]));
    host_barrier(sh, 0, world); // all handles are published
    CK(g8_mg_comm_connect(comm, sh->handles));
    host_barrier(sh, 1, world);
    int bad = 0;
    for (int fast = 0; fast < 2; ++fast) {
        bad += run_case<double>(comm, rank, world, G8_R64F, 14, fast, G8_OP_N, G8_OP_N);
        bad += run_case<double>(comm, rank, world, G8_R64F, 15, fast, G8_OP_T, G8_OP_T);
        bad += run_case<float>(comm, rank, world, G8_R32F, 6, fast, G8_OP_N, G8_OP_T);
        bad += run_case<cuDoubleComplex>(comm, rank, world, G8_C64F, 10, fast, G8_OP_N, G8_OP_N);
        bad += run_case<cuFloatComplex>(comm, rank, world, G8_C32F, 6, fast, G8_OP_C, G8_OP_T);
        // FP8 backend: local contraction into int16 residues, slab copies, shard sum mod p on the owner
        bad += run_case<double>(comm, rank, world, G8_R64F, 12, fast, G8_OP_T, G8_OP_N, G8_BACKEND_FP8);
        bad += run_case<cuFloatComplex>(comm, rank, world, G8_C32F, 7, fast, G8_OP_N, G8_OP_C, G8_BACKEND_FP8);
    }
    CK(g8_mg_comm_barrier(comm, nullptr));
    CK(cudaDeviceSynchronize());
    host_barrier(sh, 2, world);
    CK(g8_mg_comm_destroy(comm));
    sh->ok[rank] = bad == 0;
    return bad;
}

int main(int argc, char **argv) {
    // count the GPUs in a child so that the parent never initialises CUDA before forking
    int pipefd[2];
    if (pipe(pipefd)) return 2;
    if (fork() == 0) {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
        if (write(pipefd[1], &n, sizeof(n)) != sizeof(n)) _exit(1);
        _exit(0);
    }
    int ngpu = 0;
    if (read(pipefd[0], &ngpu, sizeof(ngpu)) != sizeof(ngpu)) ngpu = 0;
    wait(nullptr);
    int world = argc > 1 ? std::atoi(argv[1]) : (ngpu >= 8 ? 8 : ngpu >= 4 ? 4 : ngpu >= 2 ? 2 : 0);
    if (world < 2 || world > ngpu) {
        std::printf("mg_check SKIPPED: needs >= 2 GPUs (visible: %d)\n", ngpu);
        return 0;
    }
    Shared *sh = static_cast<Shared *>(mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
    if (sh == MAP_FAILED) return 2;
    std::memset(static_cast<void *>(sh), 0, sizeof(Shared));
    std::vector<pid_t> pids;
    for (int r = 0; r < world; ++r) {
        const pid_t p = fork();
        if (p == 0) {
            const int rc = rank_main(sh, r, world) ? 1 : 0;
            std::fflush(stdout), std::fflush(stderr); // _exit does not flush stdio
            _exit(rc);
        }
        pids.push_back(p);
    }
    int bad = 0;
    for (pid_t p : pids) {
        int st = 0;
        waitpid(p, &st, 0);
        bad += !(WIFEXITED(st) && WEXITSTATUS(st) == 0);
    }
    std::printf(bad ? "mg_check FAILED (%d ranks)\n" : "mg_check OK: %d ranks, K-sharded g8_gemm_mg == single-GPU g8_gemm\n", bad ? bad : world);
    return bad ? 1 : 0;
}
