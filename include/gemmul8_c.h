/*
 * gemmul8_b200 -- thin C ABI of the B200-native Ozaki-II GEMM emulator (the drop-in boundary).
 *
 * Plain pointers and sizes only: this is what a foreign-function binding (ctypes, cgo, JNI ...) or the
 * C++ shims in include/gemmul8.hpp bind.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference tree RIKEN-RCCS/GEMMul8 @ 71296f9, GEMMul8/...).
 *
 * All matrix / workspace pointers are DEVICE pointers; alpha / beta may be host or device pointers
 * (detected at run time exactly as reference src/inverse_scaling_real.hpp:211-213 does).  Column-major
 * BLAS conventions.  Work is enqueued on `stream`; nothing blocks the host unless timing is requested.
 *
 * Return value: 0 on success, otherwise a G8_STATUS_* / cudaError_t-compatible positive code.  The
 * reference's own API cannot report errors (it returns a timing vector and ignores CUDA status,
 * src/matmult.hpp:136-143); the C++ shims therefore drop the code, the C ABI exposes it.
 */
#ifndef GEMMUL8_C_H
#define GEMMUL8_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* gemmul8::Backend (include/gemmul8.hpp:19-20) */
enum { G8_BACKEND_INT8 = 0, G8_BACKEND_FP8 = 1 };
/* T of gemmul8::gemm<T,...> (src/gemmul8.cu:115-157) */
enum { G8_R32F = 0, G8_R64F = 1, G8_C32F = 2, G8_C64F = 3 };
/* cublasOperation_t values */
enum { G8_OP_N = 0, G8_OP_T = 1, G8_OP_C = 2 };

enum {
    G8_STATUS_SUCCESS        = 0,
    G8_STATUS_INVALID_VALUE  = 10001, /* bad enum / null pointer / num_moduli out of [2,20] (FP64) or [2,13] (FP32) / k > 2^17 (INT8) or 2^16 (FP8) */
    G8_STATUS_NOT_SUPPORTED  = 10002, /* workspace too small for the scratch a path needs (cannot happen with a g8_work_size()-sized buffer) */
    G8_STATUS_NO_DEVICE_CODE = 10003  /* the sm_100a kernels cannot run on this device: there is NO fallback path */
};

/* gemmul8::workSize<is_Complex, backend>  (include/gemmul8.hpp:25-35, src/gemmul8_real.hpp:9-47,
 * src/gemmul8_complex.hpp:9-47).  Same byte counts as the reference so caller-sized buffers keep working. */
size_t g8_work_size(int is_complex, int backend, size_t m, size_t n, size_t k, unsigned num_moduli,
                    int enable_skip_scalA, int enable_skip_scalB, size_t *workSizeA, size_t *workSizeB);

/* Argument block of one emulated GEMM: field for field the parameter list of gemmul8::gemm / gemmLt
 * (include/gemmul8.hpp:41-94); `stream` is the cudaStream_t (gemmLt's last argument, or the handle's
 * stream for gemm, src/gemmul8.cu:116-118). */
typedef struct g8_gemm_desc {
    int dtype;   /* G8_R32F .. G8_C64F */
    int backend; /* G8_BACKEND_* */
    int op_A, op_B;
    size_t m, n, k;
    const void *alpha;
    const void *A;
    size_t lda;
    const void *B;
    size_t ldb;
    const void *beta;
    void *C;
    size_t ldc;
    unsigned num_moduli;
    int fastmode;
    void *work, *workA, *workB;
    int enable_skip_scalA, enable_skip_scalB, skip_scalA, skip_scalB;
    void *stream;
} g8_gemm_desc;

/* gemmul8::gemm<T,backend> / gemmul8::gemmLt<T,backend>  (src/gemmul8.cu:115-157 -> real::gemm
 * src/gemmul8_real.hpp:53-211, complex::gemm src/gemmul8_complex.hpp:53-226).
 * phase_ns: optional double[4] {split, gemm, requant(=0: fused), crt} in nanoseconds, measured with CUDA
 * events; passing non-NULL makes the call synchronise the stream at the end (the reference always does,
 * src/common.hpp:44-57).  NULL keeps the call fully asynchronous. */
int g8_gemm(const g8_gemm_desc *d, double *phase_ns);

/* ---- stage-level entry points (used by the K-sharded multi-GPU driver and by the parity tests) ---- */

/* Stage 1 for ONE operand with externally supplied shifts: planes <- trunc(op(X) * 2^-sft) mod p_i.
 * is_A selects the m x k (A) or k x n (B) role.  Replaces scalingA/B kernels (src/scaling_fast_real.hpp:54-164). */
int g8_stage_split(int dtype, int is_A, int op, size_t rows, size_t k, const void *X, size_t ld, unsigned num_moduli,
                   int mode /*0: use sft, 1: fast (computes sft), 2: accurate bound plane + s0, 3: bound plane with the given s0*/, int16_t *sft,
                   int8_t *planes, size_t plane_stride_bytes, size_t group_stride_planes, void *stream);

/* accurate mode stage (iii): sft[i] = -(s0[i] + floor(log2P - 0.5000001*log2(cmax[i])))  (src/scaling_accu_real.hpp:6-11,157-159) */
int g8_stage_finalize_shift(int16_t *sft, const int32_t *cmax, size_t count, unsigned num_moduli, void *stream);

/* Stage 2: the low-precision GEMMs over planes (src/matmult.hpp:120-302 + src/conv_hi2mid_*.hpp).
 * epilogue (GemmEpilogue in csrc/g8_internal.cuh): 0 mod-p int8, 1 raw int32, 2 row/col max (bound GEMM), 3 complex mod-p (fused 3M tile),
 * 4 complex bound max, 5 FP8 three-piece tile, 6 FP8 bound max, 7 FP8 raw f32 (tests).
 * use_simt != 0 selects the slow dp4a cross-check kernel (tests only). */
int g8_stage_gemm(int epilogue, int use_simt, const int8_t *A_lo, size_t strideA, const int8_t *B_lo, size_t strideB, size_t m,
                  size_t n, size_t k_pad, int num_units, int first_modulus, const int *groupA, const int *groupB, void *out,
                  size_t out_stride, size_t ldc, int32_t *rowmax, int32_t *colmax, void *stream);

/* Stage 3: CRT accumulate + unscale + alpha/beta (src/inverse_scaling_real.hpp:242-278, _complex.hpp:286-326) */
int g8_stage_crt(int dtype, const void *C_mid, size_t ldmid, size_t plane_stride, size_t m, size_t n, unsigned num_moduli, void *C,
                 size_t ldc, const int16_t *sftA, const int16_t *sftB, const void *alpha, const void *beta, void *stream);

/* Stage 3 on per-shard residues (K-sharded multi-GPU, real types, INT8 backend): the residue of modulus i is
 * sym((sum_q parts[q]) mod p_i), parts `part_stride` bytes apart, each laid out like C_mid; then exactly g8_stage_crt. */
int g8_stage_crt_parts(int dtype, const void *parts, int nparts, size_t part_stride, size_t ldmid, size_t plane_stride, size_t m, size_t n,
                       unsigned num_moduli, void *C, size_t ldc, const int16_t *sftA, const int16_t *sftB, const void *alpha, const void *beta,
                       void *stream);

/* ---- K-sharded multi-GPU support (new work; the reference is single-GPU) ---- */

/* C_mid[u][col][row] = sym(C_hi[u][col][row] mod p_u) for int32 partial products that were summed across K-shards
 * (arithmetic of src/conv_hi2mid_real.hpp:9-25).  Strided so that any [col][unit][row] / [unit][col][row] layout works;
 * rows and all strides must be multiples of 4. */
int g8_stage_requant_i32(const int32_t *C_hi, size_t rows, size_t cols, size_t in_ld, size_t in_unit_stride, int num_units,
                         int first_modulus, int8_t *C_mid, size_t out_ld, size_t out_unit_stride, void *stream);

/* Residue variant: every shard reduced its partial mod p locally (int8); sum `nparts` such arrays and reduce again.
 * rows, all strides and both base pointers must be multiples of 16 (128-bit accesses). */
int g8_stage_residue_sum(const int8_t *parts, int nparts, size_t part_stride, size_t rows, size_t cols, size_t in_ld, size_t in_unit_stride,
                         int num_units, int first_modulus, int8_t *C_mid, size_t out_ld, size_t out_unit_stride, void *stream);

/* rowmax[r] = max(rowmax[r], C[r, c]), colmax[c] likewise, over an int32 slab (reduced bound product of accurate mode) */
int g8_stage_maxabs_i32(const int32_t *C, size_t rows, size_t cols, size_t ld, int32_t *rowmax, int32_t *colmax, void *stream);

/* Same, on the SUM of `nparts` int32 slabs (one per K-shard, `part_stride` elements apart): the bound product of accurate mode
 * after the fused GEMM -> scatter below. */
int g8_stage_maxabs_i32_parts(const int32_t *parts, int nparts, size_t part_stride, size_t rows, size_t cols, size_t ld, int32_t *rowmax,
                              int32_t *colmax, void *stream);

/* Bound GEMM of accurate mode over `chain` GATHERED K-slabs (K-sharded multi-GPU): plane c of A_planes (k_pad x pad256(m), K-major,
 * strideA bytes apart) and of B_planes (k_pad x n) is K-slab c of the int8 bound matrices; one accumulator sums all slabs and the
 * epilogue reduces rowmax[r] / colmax[c] with atomicMax -- the bound product is never written and no INT32 partial is exchanged. */
int g8_stage_gemm_bound_chain(const int8_t *A_planes, size_t strideA, const int8_t *B_planes, size_t strideB, size_t m, size_t n, size_t k_pad,
                              int chain, int32_t *rowmax, int32_t *colmax, void *stream);

/* Fused GEMM -> scatter over NVLink peer memory (no reference counterpart; replaces "GEMM, then NCCL all-to-all / reduce-scatter"):
 * the tcgen05 epilogue stores columns [o*n/world, (o+1)*n/world) of every unit's product directly into peer_out[o], a buffer of
 * rank o mapped with g8_peer_open (peer_out[rank] is this rank's own buffer), at column (c - o*n/world), leading dimension ldc,
 * unit stride out_stride.  epilogue 0: int8 residues mod p; 1: raw int32 partials.  n/world must be a multiple of 256.
 * peer_out is a HOST array of `world` device pointers.  The caller orders the ranks (a collective before and after). */
int g8_stage_gemm_scatter(int epilogue, const int8_t *A_lo, size_t strideA, const int8_t *B_lo, size_t strideB, size_t m, size_t n, size_t k_pad,
                          int num_units, int first_modulus, void *const *peer_out, int world, int rank, size_t out_stride, size_t ldc,
                          void *stream);

/* Device buffers mappable by the other ranks of the node: cudaMalloc + cudaIpcGetMemHandle (64-byte handle, exchanged by the host
 * side with any transport), cudaIpcOpenMemHandle / CloseMemHandle on the importing side. */
int g8_peer_alloc(size_t bytes, void **dptr, void *handle64);
int g8_peer_open(const void *handle64, void **dptr);
int g8_peer_close(void *dptr);
int g8_peer_free(void *dptr);

/* Local row statistics of op(X) (rows x k view, is_A as in g8_stage_split): amax[r] = max |x|, sumsq[r] = round-up sum of x^2 */
int g8_stage_stats(int dtype, int is_A, int op, size_t rows, size_t k, const void *X, size_t ld, double *amax, double *sumsq, void *stream);

/* Shifts from (globally reduced) statistics: kind 0 = fast-mode shift (src/scaling_fast_real.hpp:6-14, stored negated),
 * kind 1 = accurate-mode first shift s0 = 5 - ilogb(amax) (src/scaling_accu_real.hpp:39) */
int g8_stage_shift_from_stats(const double *amax, const double *sumsq, size_t count, unsigned num_moduli, int kind, int16_t *sft, void *stream);

/* ---- emulated GEMM on HOST buffers (new work; the reference's API takes device pointers only, include/gemmul8.hpp:41-94) ----
 * A plan owns the device buffers (residue planes, C_mid, staging chunks) and three streams for one problem shape; it is re-usable
 * across calls.  g8_gemm_host copies A, streams op(B) in chunks of `chunk_cols` columns (0 = 1024; rounded to 256) while the stage
 * kernels of the previous chunk run, and streams C back the same way.  hA / hB / hC are HOST pointers (pinned memory lets the copies
 * overlap; pageable memory works, slower), column-major with leading dimensions lda / ldb / ldc exactly as in g8_gemm; alpha / beta
 * are HOST scalars.  Work is ordered after everything already enqueued on `stream`, and `stream` continues after the last byte of C
 * has arrived -- the call itself does not block.  The result is bit-identical to g8_gemm on device copies of the same matrices.
 * All four types, both backends, all op_A / op_B combinations. */
typedef struct g8_host_plan g8_host_plan;
int g8_host_plan_create(g8_host_plan **plan, int dtype, int backend, int op_A, int op_B, size_t m, size_t n, size_t k, unsigned num_moduli,
                        int fastmode, size_t chunk_cols);
int g8_gemm_host(g8_host_plan *plan, const void *alpha, const void *hA, size_t lda, const void *hB, size_t ldb, const void *beta, void *hC,
                 size_t ldc, void *stream);
int g8_host_plan_destroy(g8_host_plan *plan);

/* ---- native K-sharded multi-GPU emulated GEMM (new work; the reference is single-GPU) ----
 * One process per GPU of one NVLink / NVSwitch node.  Rank r owns the K-slab  op(A)[:, K_r] (m x k_local)  and  op(B)[K_r, :]
 * (k_local x n)  and receives the column slab  C[:, r*n/world : (r+1)*n/world]  of  alpha * sum_r op(A)_r op(B)_r + beta * C.
 * All inter-rank traffic goes through peer memory (CUDA IPC) and this library's own kernels: the GEMM epilogue scatters the residue
 * tiles to their owners, small vectors (row statistics, bound maxima) are all-reduced through per-rank mailboxes in a fixed rank
 * order, ranks are ordered with a flag exchange -- no NCCL / MPI on the path.  The caller only moves one 64-byte handle per rank
 * around at start-up, with any transport it has (pipe, file, MPI, torch.distributed ...).
 *
 *   g8_mg_comm_create   allocates this rank's mailbox (sized for vectors of `max_vector_bytes`: at least 8 * (m + n) for the largest
 *                       problem) and returns its IPC handle in handle64;
 *   g8_mg_comm_connect  takes the world x 64 bytes of all ranks' handles in rank order (its own entry is ignored);
 *   g8_mg_plan_create   COLLECTIVE: allocates the planes and the peer-mapped receive area of one problem shape and exchanges the
 *                       receive-area handles through the communicator.  All four dtypes, INT8 backend, any op_A / op_B;
 *                       n / world must be a multiple of 256, world * k_local <= 2^17;
 *   g8_gemm_mg          COLLECTIVE, asynchronous on `stream`: every rank passes its slabs (device pointers, column-major, leading
 *                       dimensions lda / ldb) and gets its m x (n / world) slab of C (ld ldc); alpha / beta as in g8_gemm;
 *   g8_mg_comm_barrier  stream-ordered barrier across the ranks;  g8_mg_comm_status  != 0 after a peer timed out (~20 s).
 * Accurate mode is bit-identical to g8_gemm on the concatenated operands; fast mode is identical on all ranks and equal to the
 * single-GPU result except where the differently ordered sum of squares moves a shift across a floor() boundary. */
typedef struct g8_mg_comm g8_mg_comm;
typedef struct g8_mg_plan g8_mg_plan;
int g8_mg_comm_create(g8_mg_comm **comm, int world, int rank, size_t max_vector_bytes, void *handle64);
int g8_mg_comm_connect(g8_mg_comm *comm, const void *handles /* world x 64 bytes, rank order */);
int g8_mg_comm_barrier(g8_mg_comm *comm, void *stream);
int g8_mg_comm_status(g8_mg_comm *comm);
int g8_mg_comm_destroy(g8_mg_comm *comm);
int g8_mg_plan_create(g8_mg_plan **plan, g8_mg_comm *comm, int dtype, int op_A, int op_B, size_t m, size_t n, size_t k_local, unsigned num_moduli,
                      int fastmode);
/* the same with a backend: G8_BACKEND_FP8 contracts locally into int16 residues, moves the owners' column slabs with peer copies and sums
 * the shards mod p on the owner; world * k_local <= 2^16; all four dtypes */
int g8_mg_plan_create_backend(g8_mg_plan **plan, g8_mg_comm *comm, int dtype, int backend, int op_A, int op_B, size_t m, size_t n, size_t k_local,
                              unsigned num_moduli, int fastmode);
int g8_gemm_mg(g8_mg_plan *plan, const void *alpha, const void *A_local, size_t lda, const void *B_local, size_t ldb, const void *beta,
               void *C_slab, size_t ldc, void *stream);
int g8_mg_plan_destroy(g8_mg_plan *plan);

/* Synthetic test matrices with the reference harness' generator (testing/make_matrix.hpp:33-82):
 * element idx <- curand_init(seed, idx, 0); phi < 0: standard normal, else (u-0.5)*exp(g*phi). */
int g8_randmat(int dtype, void *X, size_t rows, size_t cols, double phi, unsigned long long seed, void *stream);

/* build / device introspection */
const char *g8_version(void);
int g8_device_supported(int device); /* 1 iff compute capability 10.0 (sm_100a cubin loads) */

#ifdef __cplusplus
}
#endif
#endif /* GEMMUL8_C_H */
