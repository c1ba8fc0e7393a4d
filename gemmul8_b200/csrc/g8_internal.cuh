// gemmul8_b200 -- internal interfaces between the stage kernels and the C-ABI orchestrator.
#pragma once
#include "g8_common.cuh"

namespace g8 {

// ---- stage 1: split -------------------------------------------------------------------------
struct SplitArgs {
    const void *X;       // operand (device), column-major
    size_t ld;           // leading dimension in elements
    size_t rows;         // rows of the "rows x inner" view (m for A, n for B)
    size_t inner;        // k
    size_t k_pad;        // pad256(k) = leading dimension of every plane
    int16_t *sft;        // shift exponents (see g8_common.cuh for the sign convention)
    int8_t *planes[3];   // real: planes[0]; complex: Re, Im, (Re+Im) mod p   [or the bound planes in mode 2]
    size_t plane_stride; // bytes between consecutive moduli
    int num_moduli;
    int row_contig;      // 1: row r contiguous along inner; 0: element (r,l) at X[l*ld + r]
    int conj;            // complex only: conjugate on load (op == C)
    int backend;         // INT8 (int8 residues) or FP8 (2-3 e4m3 pieces per residue, real types only)
};
void launch_split(const SplitArgs &a, int dtype, int mode, cudaStream_t st);
void launch_finalize_accu_shift(int16_t *sft, const int32_t *cmax, size_t count, int num_moduli, cudaStream_t st, int backend = INT8);
// K-sharded multi-GPU support
void launch_stats(const SplitArgs &a, int dtype, double *amax, double *sumsq, cudaStream_t st);
void launch_shift_from_stats(const double *amax, const double *sumsq, size_t count, int num_moduli, int kind, int16_t *sft, cudaStream_t st,
                             int backend = INT8);

// ---- stage 2: low-precision GEMMs -------------------------------------------------------------
// One "unit" is one output tile set: for unit u the kernel accumulates `nchain` products
//   acc[a] += Aplane[chainA[a][c]] (rows x K)^T-major  *  Bplane[chainB[a][c]]
// and runs the epilogue `epi` on the accumulators.
enum GemmEpilogue : int {
    EPI_MOD_I8    = 0, // C_mid[u] = int8( sym( acc0 mod p_u ) )                                   (real)
    EPI_RAW_I32   = 1, // C_hi[u]  = acc0 (int32), for the K-sharded multi-GPU partials
    EPI_BOUND_MAX = 2, // rowmax[r] = max(rowmax[r], acc0), colmax[c] likewise                     (accurate, real)
    EPI_MOD_I8_CPLX   = 3, // C_mid[u] = {sym((acc0 - acc1) mod p), sym((acc2 - acc0 - acc1) mod p)}  (complex 3M)
    EPI_BOUND_MAX_CPLX = 4, // max over max(acc0, acc1) with acc0 = ArBr + AiBi, acc1 = ArBi + AiBr
    // ---- FP8 (e4m3 x e4m3 -> f32, exact for the small-integer pieces) backend, real types ----
    EPI_F8_MOD   = 5, // three products per modulus (square moduli: AhBl, AlBh, AlBl; else Karatsuba) -> int16 residue
    EPI_F8_BOUND = 6, // bound product (one plane each), inflated by (k+1)*2^-24, row/col float maxima
    EPI_F8_RAW   = 7, // TEST ONLY: raw f32 accumulator of one product per unit
    // K-sharded multi-GPU: EPI_MOD_I8 whose residue tiles leave through shared memory + one TMA tensor store per warp (32 columns x 128 rows)
    // into the owning rank's peer-mapped buffer.  Selected internally when GemmArgs::owner_cols != 0.
    EPI_MOD_I8_SCATTER = 8,
    EPI_RAW_I32_SCATTER = 9, // same for the raw INT32 partial (the bound product of accurate mode)
    // FP8 backend, complex types: the residue planes of Re / Im / (Re+Im) go through EPI_F8_MOD once per 3M product (plane-group
    // offsets groupA[0] / groupB[0]); the bound product needs its own epilogue
    EPI_F8_BOUND_CPLX = 10, // max over max(up(|Ar||Br| + |Ai||Bi|), up(|Ar||Bi| + |Ai||Br|)), up(x) = fma_ru((2k+1) 2^-24, x, x)
    // FP8 backend, the product path: ONE piece product per unit on a full 256 x 256 tile (same operand re-use and L2 footprint as the
    // INT8 kernel); unit u -> modulus first_modulus + u / prods, product u % prods (3 piece products, x 3 plane sets for complex).
    // Output: the product's residue mod p (not wrapped) as int16; launch_f8_combine recombines (mod.hpp:106-130).
    EPI_F8_PROD = 11,
};

struct GemmArgs {
    const int8_t *A;     // base of A-side planes (k_pad x m_pad each, K-major)
    const int8_t *B;     // base of B-side planes (k_pad x n each, K-major)
    size_t strideA;      // bytes between A planes
    size_t strideB;      // bytes between B planes
    size_t m, n;         // logical extents (rows of A-side / B-side actually valid)
    size_t m_pad, k_pad; // padded extents
    int num_units;       // moduli (or 1 for the bound GEMM)
    int first_modulus;   // index of unit 0 in the moduli table
    int epi;
    // complex: plane group offsets (in planes) of Re / Im / Re+Im inside A and B
    int groupA[3], groupB[3];
    int k_true;          // FP8 bound epilogue: the unpadded k (error inflation factor)
    void *out;           // C_mid (int8 / int8x2 / int16) or C_hi (int32 / f32)
    size_t out_stride;   // elements between output planes
    size_t ldc;          // leading dimension of the output (m_pad)
    int32_t *rowmax, *colmax;
    // K-sharded multi-GPU, fused GEMM -> scatter (EPI_MOD_I8 / EPI_RAW_I32 only): see KParams in g8_gemm_i8.cu
    void *peer_out[G8_MAX_PEERS];
    size_t owner_cols; // 0 = plain single-buffer output
    int rank, world;
    // EPI_F8_PROD: products per modulus (3 real / 9 complex) and planes between the Re / Im / Re+Im plane sets;
    // EPI_MOD_I8: prods = 3 runs the three 3M products of a complex modulus as separate units (plane sets from groupA / groupB)
    int prods, set_stride;
    // EPI_BOUND_MAX: number of plane pairs summed into the accumulator (K-sharded bound product over gathered K-slabs); 0 / 1 = one
    int kchain;
};
// tcgen05 path (product).  Returns cudaError_t-compatible int.
int launch_gemm_tc(const GemmArgs &g, cudaStream_t st);
// plain dp4a path: TEST/DEBUG reference for the tensor-core kernel, never used by g8_gemm().
int launch_gemm_simt(const GemmArgs &g, cudaStream_t st);

// ---- stage 3: CRT + unscale + alpha/beta ------------------------------------------------------
struct CrtArgs {
    const void *C_mid;   // int8 (real) / int8x2 (complex), num_moduli planes
    size_t ldmid;        // m_pad
    size_t plane_stride; // elements between planes
    size_t m, n;
    int num_moduli;
    void *C;
    size_t ldc;
    const int16_t *sftA, *sftB;
    const void *alpha, *beta; // host or device pointers (resolved by the launcher)
    int backend;              // INT8: int8 residues, FP8: int16 residues and the FP8 moduli tables
    // K-sharded multi-GPU (INT8 backend): C_mid is `nparts` per-shard residue arrays `part_stride` bytes apart; the residue fed to
    // the CRT is sym((sum over parts) mod p_i).  nparts <= 1: plain C_mid.
    int nparts;
    size_t part_stride;
};
int launch_crt(const CrtArgs &c, int dtype, cudaStream_t st);

// FP8 backend: recombination of the per-product residues written by EPI_F8_PROD, laid out [modulus in batch][product][m_pad * n]:
// real (3 products): C_mid[u] = sym(recombine(c0, c1, c2) mod p) (mod.hpp:106-130);  complex (9 products): the same per 3M product,
// then {sym((rr - ii) mod p), sym((ss - rr - ii) mod p)} (conv_hi2mid_complex.hpp:130-188).  Output int16 (x 2 for complex).
// INT8 backend, complex: {re, im} = {sym((rr - ii) mod p), sym((ss - rr - ii) mod p)} (int8 x 2) from the three per-product residue
// arrays [modulus in batch][rr, ii, ss][m_pad * n] written by EPI_MOD_I8 with prods = 3 (conv_hi2mid_complex.hpp:46-127)
// `out_stride`: elements between consecutive moduli of C_mid (== elems_per_unit for the monolithic call, the full m_pad * n when the
// products cover only a column chunk)
void launch_i8_cplx_combine(const int8_t *prod, size_t elems_per_unit, int num_units, int first_modulus, int8_t *C_mid, size_t out_stride,
                            cudaStream_t st);
void launch_f8_combine(const int16_t *prod, bool cplx, size_t elems_per_unit, int num_units, int first_modulus, int16_t *C_mid, size_t out_stride,
                       cudaStream_t st);
// K-sharded complex INT8: `nparts` per-shard arrays [modulus][rr, ii, ss][elems] (part_stride BYTES apart) -> summed, reduced, recombined
void launch_i8_cplx_combine_parts(const int8_t *parts, int nparts, size_t part_stride, size_t elems_per_unit, int num_units, int first_modulus, int8_t *C_mid,
                                  size_t out_stride, cudaStream_t st);

// K-sharded FP8: `nparts` per-shard int16 residue arrays [modulus][elems] (part_stride ELEMENTS apart; complex: {re, im} interleaved,
// elems counts int16 values) -> summed over the shards and reduced to the canonical symmetric residue mod p (what f8_combine writes)
void launch_i16_sum_parts(const int16_t *parts, int nparts, size_t part_stride, size_t elems_per_unit, int num_units, int first_modulus, int16_t *C_mid,
                          size_t out_stride, cudaStream_t st);

// ---- orchestration pieces shared by g8_gemm (g8_api.cu), the host-buffer pipeline (g8_host.cu) and the multi-GPU driver (g8_mg.cu) ----
struct ContractArgs {
    bool cplx;
    int backend;
    unsigned N;
    size_t m, ncols, m_pad, k_pad;
    const int8_t *A_lo, *B_lo; // plane bases (B_lo: first column of the chunk)
    size_t sizeA, sizeB;       // bytes between planes
    size_t set_planes;         // planes between the Re / Im / Re+Im plane sets (num_mat)
    int8_t *C_mid;             // first column of the chunk
    size_t mid_plane_stride;   // ELEMENTS between C_mid planes (m_pad * n of the full matrix)
    int8_t *scratch;           // per-product residues of the multi-product paths (complex INT8, FP8)
    size_t scratch_avail;
};
int contract(const ContractArgs &c, cudaStream_t st);
int bound_gemm(bool cplx, int backend, const int8_t *A_bound, size_t sizeA, const int8_t *B_bound, size_t sizeB, size_t m, size_t ncols,
               size_t m_pad, size_t k_pad, size_t k_true, int32_t *rowmax, int32_t *colmax, cudaStream_t st);
SplitArgs make_split_args(int is_A, int op, size_t rows, size_t k, const void *X, size_t ld, unsigned N, int16_t *sft, int8_t *base,
                          size_t plane_stride, size_t group_stride_planes, int backend);
unsigned num_planes(int backend, unsigned num_moduli); // table.hpp:69-75
bool device_supported_cached();

} // namespace g8
