// gemmul8_b200 -- C++ front ends for the two entry points the reference has no counterpart for (header-only over include/gemmul8_c.h;
// nothing here adds a compiled symbol to lib/libgemmul8.{a,so}, whose symbol table stays the reference's):
//
//   gemmul8::ext::HostGemm<T, backend>   host buffers in / out: H2D, split, GEMMs, CRT and D2H pipelined over column chunks
//                                        (g8_host_plan_create / g8_gemm_host, csrc/g8_host.cu);
//   gemmul8::ext::MgComm                 one rank's end of the node-local communicator (mailboxes over CUDA IPC peer memory);
//   gemmul8::ext::MgGemm<T, backend>     K-sharded multi-GPU emulated GEMM, one process per GPU (g8_mg_plan_create_backend /
//                                        g8_gemm_mg, csrc/g8_mg.cu): rank r passes op(A)[:, K_r], op(B)[K_r, :] and receives C[:, n_r].
//
// Argument meaning follows gemmul8::gemmLt (reference include/gemmul8.hpp:68-94): column-major, cublasOperation_t ops, alpha / beta
// by pointer, num_moduli 2..20, fastmode.  Every call returns a g8 status (0 = success, include/gemmul8_c.h) instead of throwing.
#pragma once
#include "gemmul8.hpp"
#include "gemmul8_c.h"

#include <cstring>
#include <vector>

namespace gemmul8 {
namespace ext {

template <typename T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr int value = G8_R32F; };
template <> struct dtype_of<double> { static constexpr int value = G8_R64F; };
template <> struct dtype_of<cuFloatComplex> { static constexpr int value = G8_C32F; };
template <> struct dtype_of<cuDoubleComplex> { static constexpr int value = G8_C64F; };

inline int op_of(cublasOperation_t op) { return op == CUBLAS_OP_N ? G8_OP_N : op == CUBLAS_OP_T ? G8_OP_T : G8_OP_C; }
constexpr int backend_of(Backend b) { return b == Backend::INT8 ? G8_BACKEND_INT8 : G8_BACKEND_FP8; }

// ---- host buffers in / out ----
// The plan owns the device planes, the staging chunks and three streams; it is re-usable for any number of calls of its shape.
// chunk_cols: columns of op(B) / C per pipeline step (a multiple of 256; 1024 is a good default at 8192^3).
template <typename T, Backend backend = Backend::INT8> class HostGemm {
  public:
    HostGemm(cublasOperation_t op_A, cublasOperation_t op_B, size_t m, size_t n, size_t k, unsigned num_moduli, bool fastmode, size_t chunk_cols = 1024) {
        status_ = g8_host_plan_create(&plan_, dtype_of<T>::value, backend_of(backend), op_of(op_A), op_of(op_B), m, n, k, num_moduli, fastmode ? 1 : 0, chunk_cols);
    }
    ~HostGemm() {
        if (plan_) g8_host_plan_destroy(plan_);
    }
    HostGemm(const HostGemm &) = delete;
    HostGemm &operator=(const HostGemm &) = delete;
    int status() const { return status_; } // of the construction
    // hC = alpha * op(hA) * op(hB) + beta * hC; stream-ordered after `stream`, does not block the host (pinned buffers overlap best)
    int operator()(const T *alpha, const T *hA, size_t lda, const T *hB, size_t ldb, const T *beta, T *hC, size_t ldc, cudaStream_t stream = 0) {
        if (status_ != 0) return status_;
        return g8_gemm_host(plan_, alpha, hA, lda, hB, ldb, beta, hC, ldc, stream);
    }

  private:
    g8_host_plan *plan_ = nullptr;
    int status_         = 0;
};

// ---- K-sharded multi-GPU (one process per GPU of an NVLink node) ----
// Start-up: every rank constructs MgComm (the current device is its GPU), publishes handle() (64 bytes) to all other ranks with any
// transport it has, then calls connect() with the world x 64 bytes in rank order.  No NCCL / MPI is involved afterwards.
class MgComm {
  public:
    // max_vector_bytes: at least 8 * (m + n) for the largest problem this communicator will carry
    MgComm(int world, int rank, size_t max_vector_bytes) : world_(world) {
        std::memset(handle_, 0, sizeof(handle_));
        status_ = g8_mg_comm_create(&comm_, world, rank, max_vector_bytes, handle_);
    }
    ~MgComm() {
        if (comm_) g8_mg_comm_destroy(comm_);
    }
    MgComm(const MgComm &) = delete;
    MgComm &operator=(const MgComm &) = delete;
    int status() const { return status_; }
    const unsigned char *handle() const { return handle_; }                                      // this rank's 64-byte IPC handle
    int connect(const void *all_handles) { return g8_mg_comm_connect(comm_, all_handles); }      // world x 64 bytes, rank order
    int barrier(cudaStream_t stream = 0) { return g8_mg_comm_barrier(comm_, stream); }           // device-side rank barrier on `stream`
    int health() { return g8_mg_comm_status(comm_); }                                            // non-zero after a peer timed out
    int world() const { return world_; }
    g8_mg_comm *get() { return comm_; }

  private:
    g8_mg_comm *comm_ = nullptr;
    unsigned char handle_[64];
    int world_, status_ = 0;
};

// COLLECTIVE construction and calls: every rank of the communicator makes them in the same order.  n / world must be a multiple of
// 256; world * k_local <= 2^17 (INT8) / 2^16 (FP8).  Accurate mode is bit-identical to gemmLt on the concatenated operands.
template <typename T, Backend backend = Backend::INT8> class MgGemm {
  public:
    MgGemm(MgComm &comm, cublasOperation_t op_A, cublasOperation_t op_B, size_t m, size_t n, size_t k_local, unsigned num_moduli, bool fastmode) {
        status_ = g8_mg_plan_create_backend(&plan_, comm.get(), dtype_of<T>::value, backend_of(backend), op_of(op_A), op_of(op_B), m, n, k_local, num_moduli,
                                            fastmode ? 1 : 0);
    }
    ~MgGemm() {
        if (plan_) g8_mg_plan_destroy(plan_); // collective too: it waits for the peers before the receive area is freed
    }
    MgGemm(const MgGemm &) = delete;
    MgGemm &operator=(const MgGemm &) = delete;
    int status() const { return status_; }
    // C_slab (m x n / world, ld ldc) = alpha * op(A)[:, K_r] ... summed over the ranks ... + beta * C_slab; device pointers, asynchronous
    int operator()(const T *alpha, const T *A_local, size_t lda, const T *B_local, size_t ldb, const T *beta, T *C_slab, size_t ldc, cudaStream_t stream = 0) {
        if (status_ != 0) return status_;
        return g8_gemm_mg(plan_, alpha, A_local, lda, B_local, ldb, beta, C_slab, ldc, stream);
    }

  private:
    g8_mg_plan *plan_ = nullptr;
    int status_       = 0;
};

} // namespace ext
} // namespace gemmul8
