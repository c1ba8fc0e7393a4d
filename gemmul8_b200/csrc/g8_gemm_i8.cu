// gemmul8_b200 -- stage 2: the num_moduli exact INT8 x INT8 -> INT32 contractions on tcgen05 tensor cores,
// with the modular requantisation fused into the TMEM epilogue.
//
// Replaces the reference's library calls (matmult.hpp:120-302: cublasGemmEx / cublasLtMatmul, one per
// modulus) AND its per-modulus conv_hi2mid pass (conv_hi2mid_real.hpp:9-25, conv_hi2mid_complex.hpp:46-127),
// AND the row/column max passes over the bound product in accurate mode (scaling_accu_real.hpp:142-226).
// No cuBLAS / cuBLASLt / CUTLASS on this path: TMA (cp.async.bulk.tensor, 128B swizzle) -> smem ring ->
// tcgen05.mma.kind::i8 issued by one thread -> s32 accumulators in TMEM -> tcgen05.ld -> epilogue.
//
// Orientation.  Both operand sets are K-major (row r of op(A) / column c of op(B) is k_pad contiguous
// bytes), which is exactly the canonical K-major SWIZZLE_128B UMMA layout.  We compute D = B_lo^T-tile x
// A_lo-tile, i.e. the MMA "M" side (TMEM lanes) walks columns c of C and the MMA "N" side (TMEM columns)
// walks rows r of C, so that after tcgen05.ld.32x32b every thread holds a run of CONSECUTIVE ROWS of one
// column of the column-major output and can store 16 bytes at a time.
//
// One persistent CTA per SM, by default grouped into CTA pairs (cta_group::2, one 256 x 256 tile per pair); work = (unit, tile)
// pairs in unit-major order so that all SMs work on the same plane pair (it stays L2-resident); the TMEM accumulator is a ring
// of slots so the epilogue of tile t overlaps the MMAs of tile t+1.  A "unit" is ONE low-precision product: a modulus (real INT8),
// one of the three 3M products of a modulus (complex INT8) or one FP8 piece product (kind::f8f6f4); multi-product residues are
// recombined by a small pass in g8_api.cu.  The *_SCATTER epilogues are the fused GEMM -> NVLink exchange of the K-sharded
// multi-GPU path (residue tiles leave through swizzled shared memory + one TMA tensor store per warp into the owning rank's peer-mapped buffer).
#include "g8_internal.cuh"

#include <cuda.h> // CUtensorMap (types only; the encoder is fetched through the runtime, no -lcuda)
#include <atomic>
#include <cstdlib>
#include <mutex>

namespace g8 {

// ------------------------------------------------------------------------------------------------
// raw PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// the same loads with an L2 eviction-priority hint (createpolicy): the band operand that is re-used across a whole sweep is kept
// (evict_last), the streamed panels / the output must not displace it
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_3d_hint(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm_hint(void *smem_dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1, int c2, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], s8 x s8 -> s32, issued by ONE thread
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// e4m3 x e4m3 -> f32 (kind::f8f6f4)
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- cta_group::2 (CTA pair) variants ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `smem_addr` (a shared::cta address of THIS CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair into its OWN smem, completing on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_3d_2sm(void *smem_dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all prior MMAs -> arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------------------
// K-major, SWIZZLE_128B canonical layout: rows of 128 bytes, 8-row groups 1024 B apart.
//   bits [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1) | [32,46) SBO >> 4 (= 64)
//   bits [46,48) descriptor version (1 on sm_100) | [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
           (uint64_t(2) << 61);
}
// instruction descriptor, kind::i8: c_format = S32 (2) [4,6); a/b format = signed 8-bit (1) [7,10)/[10,13);
// a/b major = K (0); N >> 3 at [17,23); M >> 4 at [24,29); no saturation (sums are < 2^31 by construction)
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f8f6f4: c_format = F32 (1); a/b format = E4M3 (0); K-major both
__host__ __device__ constexpr uint32_t make_idesc_f8(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// kernel configuration
// ------------------------------------------------------------------------------------------------
constexpr int BLOCK_K   = 128; // bytes of K per pipeline stage (one 128B swizzle row)
constexpr int UMMA_K    = 32;  // K per tcgen05.mma for 8-bit operands
constexpr int TILE_LANE = 128; // MMA M: columns of C handled per tile (TMEM lanes)
constexpr int NUM_THREADS = 192; // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue

template <int EPI> struct EpiCfg;
template <> struct EpiCfg<EPI_MOD_I8>         { static constexpr int TILE_COL = 256, NACC = 1, NCHAIN = 1; };
template <> struct EpiCfg<EPI_RAW_I32>        { static constexpr int TILE_COL = 256, NACC = 1, NCHAIN = 1; };
template <> struct EpiCfg<EPI_BOUND_MAX>      { static constexpr int TILE_COL = 256, NACC = 1, NCHAIN = 1; };
template <> struct EpiCfg<EPI_MOD_I8_CPLX>    { static constexpr int TILE_COL = 128, NACC = 3, NCHAIN = 1; };
template <> struct EpiCfg<EPI_BOUND_MAX_CPLX> { static constexpr int TILE_COL = 128, NACC = 2, NCHAIN = 2; };
template <> struct EpiCfg<EPI_F8_MOD>         { static constexpr int TILE_COL = 128, NACC = 3, NCHAIN = 1; };
template <> struct EpiCfg<EPI_F8_BOUND>       { static constexpr int TILE_COL = 256, NACC = 1, NCHAIN = 1; };
template <> struct EpiCfg<EPI_F8_RAW>         { static constexpr int TILE_COL = 256, NACC = 1, NCHAIN = 1; };
template <> struct EpiCfg<EPI_MOD_I8_SCATTER> { static constexpr int TILE_COL = 256, NACC = 1, NCHAIN = 1; };
template <> struct EpiCfg<EPI_RAW_I32_SCATTER> { static constexpr int TILE_COL = 256, NACC = 1, NCHAIN = 1; };
template <> struct EpiCfg<EPI_F8_BOUND_CPLX>  { static constexpr int TILE_COL = 128, NACC = 2, NCHAIN = 2; };
template <> struct EpiCfg<EPI_F8_PROD>        { static constexpr int TILE_COL = 256, NACC = 1, NCHAIN = 1; };
// Scatter staging.
// EPI_MOD_I8_SCATTER: per epilogue warp G8_SCAT_BUFS buffers of 32 columns x 128 rows (4 KB, 128 B per column, 16-byte chunks XOR-swizzled
// by column & 7 = CU_TENSOR_MAP_SWIZZLE_128B, which also makes the 16-byte shared stores of the 32 lanes conflict-free); ONE TMA tensor
// store per buffer.  (Before: one 256-byte cp.async.bulk per THREAD, 256 copy-engine operations per CTA and tile against 8 now; measured neutral
// within noise on 1 and 8 GPUs, profiles/r02i_*.)
// EPI_RAW_I32_SCATTER (non-default INT32 bound exchange): per thread one 128-byte bulk copy per TMEM chunk from a padded row.
#ifndef G8_SCAT_BUFS
#define G8_SCAT_BUFS 2
#endif
constexpr int SCAT_BUF_BYTES = 32 * 128, SCAT_TMA_WARP_BYTES = G8_SCAT_BUFS * SCAT_BUF_BYTES, SCAT_TMA_BYTES = 4 * SCAT_TMA_WARP_BYTES;
constexpr int SCAT_PITCH = 128 + 16, SCAT_WARP_BYTES = 32 * SCAT_PITCH, SCAT_BYTES = 4 * SCAT_WARP_BYTES;
template <int EPI> constexpr bool is_f8 = (EPI == EPI_F8_MOD || EPI == EPI_F8_BOUND || EPI == EPI_F8_RAW || EPI == EPI_F8_BOUND_CPLX || EPI == EPI_F8_PROD);

// CG = 1: one CTA per tile (128 columns of C x TILE_COL rows).  CG = 2: a CTA pair (tcgen05 cta_group::2) shares a
// 256-column tile; each CTA stages its own 128 columns' operand plus HALF of the row-side operand, which halves the
// shared-memory traffic per MMA and deepens the TMA ring.
template <int EPI, int CG = 1> struct KernelShape {
    using C = EpiCfg<EPI>;
    static constexpr int TILE_COL   = C::TILE_COL;                 // MMA N: rows of C per tile (TMEM columns)
    static constexpr int STAGE_L    = TILE_LANE * BLOCK_K;         // bytes: lane-side operand (B_lo tile), per CTA
    static constexpr int STAGE_C    = TILE_COL / CG * BLOCK_K;     // bytes: column-side operand (A_lo tile), per CTA
    static constexpr int STAGE      = STAGE_L + STAGE_C;
    static constexpr int SCAT       = EPI == EPI_MOD_I8_SCATTER ? SCAT_TMA_BYTES : EPI == EPI_RAW_I32_SCATTER ? SCAT_BYTES : 0; // epilogue staging
    static constexpr int BARS       = EPI == EPI_MOD_I8_SCATTER ? 1024 : 256; // barrier area (keeps the swizzled staging 1024-byte aligned)
    static constexpr int NUM_STAGES = (220 * 1024 - SCAT) / STAGE > 8 ? 8 : (220 * 1024 - SCAT) / STAGE;
    // TMEM is a ring of accumulator SLOTS of TILE_COL columns; a tile takes NACC consecutive slots.  With more slots than
    // NACC (2 vs 1, 4 vs 3) the MMAs of the next tile start while the epilogue still drains the previous one.
    static constexpr int NUM_BUF    = 512 / TILE_COL;
    static constexpr int SMEM_BYTES = NUM_STAGES * STAGE + 1024 /*align*/ + BARS + SCAT;
};

struct TileCoord {
    int unit, tl, tc; // unit, lane-side tile index (along n), column-side tile index (along m)
};

// unit-major; inside a unit, bands of GROUP lane-tiles are swept along the column-tile direction so that
// the CTAs running concurrently share a small set of operand panels in L2.
__device__ __forceinline__ TileCoord tile_coord(int t, int tiles_l, int tiles_c, int tl_rot = 0, int GROUP = 16) {
    const int per_unit  = tiles_l * tiles_c;
    TileCoord r;
    r.unit        = t / per_unit;
    int idx       = t - r.unit * per_unit;
    const int band_sz = GROUP * tiles_c;
    const int band    = idx / band_sz;
    idx -= band * band_sz;
    const int gl = min(GROUP, tiles_l - band * GROUP);
    r.tc         = idx / gl;
    r.tl         = band * GROUP + (idx - r.tc * gl);
    r.tl += tl_rot;
    if (r.tl >= tiles_l) r.tl -= tiles_l;
    return r;
}

struct KParams {
    int tiles_l, tiles_c, num_units, first_modulus, kblocks;
    int n, m; // valid extents
    int groupA[3], groupB[3];
    void *out;
    size_t out_stride, ldc;
    int32_t *rowmax, *colmax;
    float inflate; // FP8 bound: (k + 1) * 2^-24
    // fused GEMM -> scatter of the K-sharded multi-GPU path: columns [o * owner_cols, (o + 1) * owner_cols) of the output go to
    // peer_out[o] (a peer-mapped buffer of rank o, or our own), at column (c - o * owner_cols).  owner_cols == 0: plain output `out`.
    void *peer_out[G8_MAX_PEERS];
    int owner_cols;
    int prods, set_stride; // EPI_F8_PROD
    int l2hint; // bit 0: evict_last for the band (lane-side) operand loads; bit 1: evict_first for the column-side panels; bit 2: streaming C_mid stores
    int kchain; // bound epilogues: the accumulator sums `kchain` plane pairs (plane c of both operands = K-slab c of a K-sharded bound product)
    int group;  // lane tiles per rasterisation band (tile_coord)
    int tl_rot; // rotation of the lane-tile sweep so that the ranks do not all target the same owner at the same time
};

// FP8 plane bookkeeping (table.hpp:69-75): modulus idx owns planes [base, base + 2) (square moduli, idx < 6) or [base, base + 3)
__host__ __device__ __forceinline__ int f8_plane_base(int idx) { return idx < 6 ? 2 * idx : 12 + 3 * (idx - 6); }

// (acc, chain) -> (A group, B group)
template <int EPI> __device__ __forceinline__ void chain_groups(int acc, int c, int &ga, int &gb) {
    if constexpr (EPI == EPI_MOD_I8_CPLX) {
        ga = acc, gb = acc; // ArBr, AiBi, (Ar+Ai)(Br+Bi)
    } else if constexpr (EPI == EPI_BOUND_MAX_CPLX || EPI == EPI_F8_BOUND_CPLX) {
        ga = c;                      // acc0 = |Ar||Br| + |Ai||Bi| ; acc1 = |Ar||Bi| + |Ai||Br|
        gb = (acc == 0) ? c : 1 - c;
    } else {
        ga = 0, gb = 0;
    }
}

// EPI_MOD_I8_SCATTER: one store tensor map per owner, view {ldc rows (bytes), owner_cols, units} of its receive area
struct PeerMaps {
    CUtensorMap m[G8_MAX_PEERS];
};

// products chained into one accumulator: the bound epilogues sum `kchain` gathered K-slabs (K-sharded multi-GPU; 1 in the single-GPU call)
template <int EPI> __device__ __forceinline__ int chain_len(int kchain) {
    if constexpr (EPI == EPI_BOUND_MAX || EPI == EPI_F8_BOUND) return kchain;
    else if constexpr (EPI == EPI_BOUND_MAX_CPLX || EPI == EPI_F8_BOUND_CPLX) return 2 * kchain;
    else return EpiCfg<EPI>::NCHAIN;
}

template <int EPI, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_i8_tc_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapC, const KParams P,
                  const __grid_constant__ PeerMaps PM) {
    using KS = KernelShape<EPI, CG>;
    using EC = EpiCfg<EPI>;
    constexpr int TILE_COL = KS::TILE_COL, NUM_STAGES = KS::NUM_STAGES, NUM_BUF = KS::NUM_BUF;
    // CTA pair bookkeeping: rank 0 (leader) issues the MMAs and owns the "full" / "TMEM empty" barriers of the pair
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader   = rank == 0;
    const int cid = blockIdx.x / CG, ncl = gridDim.x / CG; // tile-scheduler identity: one work stream per CTA (pair)

    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t *bars      = reinterpret_cast<uint64_t *>(smem + NUM_STAGES * KS::STAGE);
    uint64_t *full_bar  = bars;                    // [NUM_STAGES] TMA -> MMA
    uint64_t *empty_bar = bars + NUM_STAGES;       // [NUM_STAGES] MMA -> TMA
    uint64_t *tfull_bar = bars + 2 * NUM_STAGES;   // [NUM_BUF]    MMA -> epilogue
    uint64_t *tempty_bar = tfull_bar + NUM_BUF;    // [NUM_BUF]    epilogue -> MMA
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + NUM_BUF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = P.num_units * P.tiles_l * P.tiles_c;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapL);
        tma_prefetch_desc(&mapC);
        for (int s = 0; s < NUM_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < NUM_BUF; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], 4 * CG); // one arrive per epilogue warp (of both CTAs of a pair)
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); // the peer's barriers must be initialised before any remote arrive / TMA completion
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint64_t polL = l2_policy_evict_last(), polC = l2_policy_evict_first();
            const bool hintL = P.l2hint & 1, hintC = P.l2hint & 2;
            for (int t = cid; t < total_tiles; t += ncl) {
                const TileCoord tc = tile_coord(t, P.tiles_l, P.tiles_c, P.tl_rot, P.group);
                const int nchain = chain_len<EPI>(P.kchain);
                for (int acc = 0; acc < EC::NACC; ++acc)
                    for (int c = 0; c < nchain; ++c) {
                        int planeA, planeB;
                        if constexpr (EPI == EPI_F8_PROD) {
                            // unit -> (modulus, product); product q = 3 * (plane set: Re, Im, Re+Im) + piece product
                            const int idx = P.first_modulus + tc.unit / P.prods, q = tc.unit % P.prods;
                            const int set = q / 3, a = q - 3 * set, base = f8_plane_base(idx) + set * P.set_stride;
                            const bool sq = idx < 6;
                            planeA = base + (sq ? (a == 0 ? 0 : 1) : a);
                            planeB = base + (sq ? (a == 1 ? 0 : 1) : a);
                        } else if constexpr (EPI == EPI_F8_MOD) {
                            // square moduli: AhBl, AlBh, AlBl (gemmul8_real.hpp:159-170); otherwise Karatsuba hi*hi, lo*lo, sum*sum (:171-180)
                            const int idx = P.first_modulus + tc.unit, base = f8_plane_base(idx);
                            const bool sq = idx < 6;
                            planeA = P.groupA[0] + base + (sq ? (acc == 0 ? 0 : 1) : acc); // group offset: 0 (real) or the Re / Im / Re+Im set
                            planeB = P.groupB[0] + base + (sq ? (acc == 1 ? 0 : 1) : acc);
                        } else if constexpr (EPI == EPI_MOD_I8 || EPI == EPI_MOD_I8_SCATTER) {
                            // real: one product per modulus.  complex (prods = 3): unit -> (modulus, 3M product q) with the plane sets
                            // Re / Im / Re+Im of BOTH operands selected by q
                            const int q = tc.unit % P.prods, mu = tc.unit / P.prods;
                            planeA = P.groupA[q] + mu, planeB = P.groupB[q] + mu;
                        } else if constexpr (EPI == EPI_BOUND_MAX || EPI == EPI_F8_BOUND) {
                            planeA = P.groupA[0] + tc.unit + c, planeB = P.groupB[0] + tc.unit + c; // chained K-slabs (c = 0 only in the single-GPU call)
                        } else if constexpr (EPI == EPI_BOUND_MAX_CPLX || EPI == EPI_F8_BOUND_CPLX) {
                            // K-slab c / 2 of the gathered planes [slab][|Re|, |Im|]; inside a slab the two products of chain_groups
                            int ga, gb;
                            chain_groups<EPI>(acc, c & 1, ga, gb);
                            planeA = P.groupA[ga] + tc.unit + 2 * (c >> 1), planeB = P.groupB[gb] + tc.unit + 2 * (c >> 1);
                        } else {
                            int ga, gb;
                            chain_groups<EPI>(acc, c, ga, gb);
                            planeA = P.groupA[ga] + tc.unit, planeB = P.groupB[gb] + tc.unit;
                        }
                        for (int kb = 0; kb < P.kblocks; ++kb) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            unsigned char *sL = smem + stage * KS::STAGE;
                            unsigned char *sC = sL + KS::STAGE_L;
                            if constexpr (CG == 2) {
                                // both CTAs load their share; all bytes complete on the leader's barrier
                                if (leader) mbar_expect_tx(&full_bar[stage], KS::STAGE * 2);
                                const uint32_t lbar = mapa(smem_u32(&full_bar[stage]), 0);
                                const int cl = tc.tl * (TILE_LANE * 2) + (int)rank * TILE_LANE, cc = tc.tc * TILE_COL + (int)rank * (TILE_COL / 2);
                                if (hintL) tma_load_3d_2sm_hint(sL, &mapL, lbar, kb * BLOCK_K, cl, planeB, polL);
                                else tma_load_3d_2sm(sL, &mapL, lbar, kb * BLOCK_K, cl, planeB);
                                if (hintC) tma_load_3d_2sm_hint(sC, &mapC, lbar, kb * BLOCK_K, cc, planeA, polC);
                                else tma_load_3d_2sm(sC, &mapC, lbar, kb * BLOCK_K, cc, planeA);
                            } else {
                                mbar_expect_tx(&full_bar[stage], KS::STAGE);
                                if (hintL) tma_load_3d_hint(sL, &mapL, &full_bar[stage], kb * BLOCK_K, tc.tl * TILE_LANE, planeB, polL);
                                else tma_load_3d(sL, &mapL, &full_bar[stage], kb * BLOCK_K, tc.tl * TILE_LANE, planeB);
                                if (hintC) tma_load_3d_hint(sC, &mapC, &full_bar[stage], kb * BLOCK_K, tc.tc * TILE_COL, planeA, polC);
                                else tma_load_3d(sC, &mapC, &full_bar[stage], kb * BLOCK_K, tc.tc * TILE_COL, planeA);
                            }
                            if (++stage == NUM_STAGES) stage = 0, phase ^= 1;
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = is_f8<EPI> ? make_idesc_f8(TILE_LANE * CG, TILE_COL) : make_idesc_i8(TILE_LANE * CG, TILE_COL);
            int stage = 0, buf = 0;
            uint32_t phase = 0, tphase = 0;
            for (int t = cid; t < total_tiles; t += ncl) {
                for (int acc = 0; acc < EC::NACC; ++acc) {
                    mbar_wait(&tempty_bar[buf], tphase ^ 1); // slot drained by the epilogue of an earlier tile
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * TILE_COL;
                    uint32_t accumulate   = 0;
                    const int nchain      = chain_len<EPI>(P.kchain);
                    for (int c = 0; c < nchain; ++c)
                        for (int kb = 0; kb < P.kblocks; ++kb) {
                            mbar_wait(&full_bar[stage], phase);
                            tc_fence_after();
                            const uint32_t sL = smem_u32(smem + stage * KS::STAGE);
                            const uint64_t dL = make_smem_desc(sL), dC = make_smem_desc(sL + KS::STAGE_L);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                const uint64_t a = dL + (uint64_t)(k * UMMA_K >> 4), b = dC + (uint64_t)(k * UMMA_K >> 4);
                                if constexpr (CG == 2) {
                                    if constexpr (is_f8<EPI>) umma_f8_2sm(d_tmem, a, b, idesc, accumulate);
                                    else umma_i8_2sm(d_tmem, a, b, idesc, accumulate);
                                } else {
                                    if constexpr (is_f8<EPI>) umma_f8(d_tmem, a, b, idesc, accumulate);
                                    else umma_i8(d_tmem, a, b, idesc, accumulate);
                                }
                                accumulate = 1;
                            }
                            // frees the smem slot (in both CTAs of a pair) once these MMAs retire
                            if constexpr (CG == 2) umma_commit_2sm(&empty_bar[stage]);
                            else umma_commit(&empty_bar[stage]);
                            if (++stage == NUM_STAGES) stage = 0, phase ^= 1;
                        }
                    // this accumulator is complete
                    if constexpr (CG == 2) umma_commit_2sm(&tfull_bar[buf]);
                    else umma_commit(&tfull_bar[buf]);
                    if (++buf == NUM_BUF) buf = 0, tphase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
        const int q = warp & 3;
        int buf = 0;
        [[maybe_unused]] int scat_buf = 0; // EPI_MOD_I8_SCATTER: staging buffer of this warp the next tensor store leaves from
        uint32_t tphase = 0;
        for (int t = cid; t < total_tiles; t += ncl) {
            const TileCoord tc = tile_coord(t, P.tiles_l, P.tiles_c, P.tl_rot, P.group);
            // the NACC accumulators of this tile sit in consecutive ring slots
            uint32_t ta[3] = {0, 0, 0};
            int sb = buf;
            uint32_t sp = tphase;
#pragma unroll
            for (int a = 0; a < EC::NACC; ++a) {
                mbar_wait(&tfull_bar[sb], sp);
                ta[a] = tmem_base + ((uint32_t)(q * 32) << 16) + sb * TILE_COL;
                if (++sb == NUM_BUF) sb = 0, sp ^= 1;
            }
            tc_fence_after();
            const uint32_t taddr0 = ta[0], ta1 = ta[1], ta2 = ta[2];
            (void)ta1, (void)ta2;
            const int col_c   = tc.tl * (TILE_LANE * CG) + (int)rank * TILE_LANE + q * 32 + lane; // column of C owned by this thread
            const int row0    = tc.tc * TILE_COL;                  // first row of C of this tile
            const bool col_ok = col_c < P.n;
            const int midx    = P.first_modulus + ((EPI == EPI_MOD_I8 || EPI == EPI_MOD_I8_SCATTER) ? tc.unit / P.prods : tc.unit);
            // output base and column index inside it (peer scatter: the whole 256-column tile belongs to one owner)
            char *out_base = static_cast<char *>(P.out);
            int col_o      = col_c;
            if (P.owner_cols) {
                const int owner = (tc.tl * (TILE_LANE * CG)) / P.owner_cols;
                out_base        = static_cast<char *>(P.peer_out[owner]);
                col_o           = col_c - owner * P.owner_cols;
            }

            if constexpr (EPI == EPI_MOD_I8) {
                const int32_t p = g8d_moduli[INT8][midx], pinv = g8d_pinv32[INT8][midx];
                int8_t *dst = reinterpret_cast<int8_t *>(out_base) + (size_t)tc.unit * P.out_stride + (size_t)col_o * P.ldc + row0;
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_COL; c0 += 32) {
                    int32_t v[32];
                    tmem_ld32(taddr0 + c0, v);
                    tmem_ld_wait();
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int32_t r0 = mod_i32(v[4 * j], p, pinv), r1 = mod_i32(v[4 * j + 1], p, pinv);
                        const int32_t r2 = mod_i32(v[4 * j + 2], p, pinv), r3 = mod_i32(v[4 * j + 3], p, pinv);
                        w[j] = (uint32_t)(r0 & 0xFF) | ((uint32_t)(r1 & 0xFF) << 8) | ((uint32_t)(r2 & 0xFF) << 16) | ((uint32_t)r3 << 24);
                    }
                    if (col_ok) {
                        if (P.l2hint & 4) { // C_mid is written once and read much later (CRT): do not let it displace the operand band
                            __stcs(reinterpret_cast<uint4 *>(dst + c0), make_uint4(w[0], w[1], w[2], w[3]));
                            __stcs(reinterpret_cast<uint4 *>(dst + c0 + 16), make_uint4(w[4], w[5], w[6], w[7]));
                        } else {
                            *reinterpret_cast<uint4 *>(dst + c0)      = make_uint4(w[0], w[1], w[2], w[3]);
                            *reinterpret_cast<uint4 *>(dst + c0 + 16) = make_uint4(w[4], w[5], w[6], w[7]);
                        }
                    }
                }
            } else if constexpr (EPI == EPI_MOD_I8_SCATTER) {
                // residues -> swizzled shared staging (32 columns x 128 rows per warp) -> ONE TMA tensor store per warp and 128 rows into
                // the owner's receive area: NVLink sees 128-byte row segments issued by the copy engine, the epilogue warps never wait
                // on remote latency, and the SM's copy engine handles 8 store operations per tile instead of 256.
                const int32_t p = g8d_moduli[INT8][midx], pinv = g8d_pinv32[INT8][midx];
                const int owner  = (tc.tl * (TILE_LANE * CG)) / P.owner_cols;
                const int col_w  = col_o - lane; // first of this warp's 32 columns inside the owner's slab
                unsigned char *stg_w = smem + NUM_STAGES * KS::STAGE + KS::BARS + (warp - 2) * SCAT_TMA_WARP_BYTES;
                const int sw = (lane & 7) << 4;
#pragma unroll 1
                for (int h0 = 0; h0 < TILE_COL; h0 += 128) {
                    // the copy engine must have finished READING the buffer we are about to overwrite
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(G8_SCAT_BUFS - 1) : "memory");
                    __syncwarp();
                    unsigned char *buf_s = stg_w + scat_buf * SCAT_BUF_BYTES;
                    unsigned char *row_s = buf_s + lane * 128;
#pragma unroll 1
                    for (int c0 = 0; c0 < 128; c0 += 32) {
                        int32_t v[32];
                        tmem_ld32(taddr0 + h0 + c0, v);
                        tmem_ld_wait();
                        uint32_t w[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int32_t r0 = mod_i32(v[4 * j], p, pinv), r1 = mod_i32(v[4 * j + 1], p, pinv);
                            const int32_t r2 = mod_i32(v[4 * j + 2], p, pinv), r3 = mod_i32(v[4 * j + 3], p, pinv);
                            w[j] = (uint32_t)(r0 & 0xFF) | ((uint32_t)(r1 & 0xFF) << 8) | ((uint32_t)(r2 & 0xFF) << 16) | ((uint32_t)r3 << 24);
                        }
                        *reinterpret_cast<uint4 *>(row_s + (c0 ^ sw))        = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4 *>(row_s + ((c0 + 16) ^ sw)) = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                    fence_proxy_async(); // generic-proxy writes of this thread -> visible to the async (copy engine) proxy
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&PM.m[owner]),
                                     "r"(row0 + h0), "r"(col_w), "r"(tc.unit), "r"(smem_u32(buf_s))
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    scat_buf = (scat_buf + 1 == G8_SCAT_BUFS) ? 0 : scat_buf + 1;
                }
            } else if constexpr (EPI == EPI_RAW_I32_SCATTER) {
                // 32 rows x 4 B = 128 B per column and TMEM chunk: same staging row, one bulk copy per chunk
                int32_t *dst = reinterpret_cast<int32_t *>(out_base) + (size_t)tc.unit * P.out_stride + (size_t)col_o * P.ldc + row0;
                unsigned char *stg = smem + NUM_STAGES * KS::STAGE + 256 + (warp - 2) * SCAT_WARP_BYTES + lane * SCAT_PITCH;
                const uint32_t stg_s = smem_u32(stg);
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_COL; c0 += 32) {
                    int32_t v[32];
                    tmem_ld32(taddr0 + c0, v);
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<int4 *>(stg + 16 * j) = make_int4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    fence_proxy_async();
                    if (col_ok)
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c0), "r"(stg_s), "r"(128) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else if constexpr (EPI == EPI_RAW_I32) {
                int32_t *dst = reinterpret_cast<int32_t *>(out_base) + (size_t)tc.unit * P.out_stride + (size_t)col_o * P.ldc + row0;
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_COL; c0 += 32) {
                    int32_t v[32];
                    tmem_ld32(taddr0 + c0, v);
                    tmem_ld_wait();
                    if (col_ok) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<int4 *>(dst + c0 + 4 * j) = make_int4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            } else if constexpr (EPI == EPI_BOUND_MAX) {
                int32_t cmax = 0;
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_COL; c0 += 32) {
                    int32_t v[32];
                    tmem_ld32(taddr0 + c0, v);
                    tmem_ld_wait();
                    int32_t mine = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        cmax            = max(cmax, v[j]);
                        const int32_t r = __reduce_max_sync(0xffffffffu, v[j]); // max over the 32 columns of C of this warp
                        mine            = (j == lane) ? r : mine;
                    }
                    if (mine > 0) atomicMax(&P.rowmax[row0 + c0 + lane], mine);
                }
                if (col_ok && cmax > 0) atomicMax(&P.colmax[col_c], cmax);
            } else if constexpr (EPI == EPI_MOD_I8_CPLX) {
                const int32_t p = g8d_moduli[INT8][midx], pinv = g8d_pinv32[INT8][midx];
                int8_t *dst = reinterpret_cast<int8_t *>(P.out) + ((size_t)tc.unit * P.out_stride + (size_t)col_c * P.ldc + row0) * 2;
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_COL; c0 += 16) {
                    int32_t a0[16], a1[16], a2[16];
                    tmem_ld16(taddr0 + c0, a0);
                    tmem_ld16(ta1 + c0, a1);
                    tmem_ld16(ta2 + c0, a2);
                    tmem_ld_wait();
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        int32_t o[4];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            // reduce each product first (k up to 2^17 makes the raw differences overflow int32)
                            const int32_t x0 = a0[2 * j + e] - p * __mulhi(a0[2 * j + e], pinv);
                            const int32_t x1 = a1[2 * j + e] - p * __mulhi(a1[2 * j + e], pinv);
                            const int32_t x2 = a2[2 * j + e] - p * __mulhi(a2[2 * j + e], pinv);
                            o[2 * e]     = mod_i32(x0 - x1, p, pinv);      // Re = ArBr - AiBi
                            o[2 * e + 1] = mod_i32(x2 - x0 - x1, p, pinv); // Im = (Ar+Ai)(Br+Bi) - ArBr - AiBi
                        }
                        w[j] = (uint32_t)(o[0] & 0xFF) | ((uint32_t)(o[1] & 0xFF) << 8) | ((uint32_t)(o[2] & 0xFF) << 16) | ((uint32_t)o[3] << 24);
                    }
                    if (col_ok) {
                        *reinterpret_cast<uint4 *>(dst + 2 * c0)      = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4 *>(dst + 2 * c0 + 16) = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                }
            } else if constexpr (EPI == EPI_F8_RAW) {
                int32_t *dst = reinterpret_cast<int32_t *>(P.out) + (size_t)tc.unit * P.out_stride + (size_t)col_c * P.ldc + row0;
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_COL; c0 += 32) {
                    int32_t v[32];
                    tmem_ld32(taddr0 + c0, v);
                    tmem_ld_wait();
                    if (col_ok) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<int4 *>(dst + c0 + 4 * j) = make_int4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            } else if constexpr (EPI == EPI_F8_BOUND) {
                // find_max.hpp:82-96,166-188: max of fma_ru(ku, v, v); non-negative floats order like their bit patterns
                int32_t cmax = 0;
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_COL; c0 += 32) {
                    int32_t v[32];
                    tmem_ld32(taddr0 + c0, v);
                    tmem_ld_wait();
                    int32_t mine = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float f   = __int_as_float(v[j]);
                        const int32_t x = __float_as_int(fmaxf(__fmaf_ru(P.inflate, f, f), 0.0f));
                        cmax            = max(cmax, x);
                        const int32_t r = __reduce_max_sync(0xffffffffu, x);
                        mine            = (j == lane) ? r : mine;
                    }
                    if (mine > 0) atomicMax(&P.rowmax[row0 + c0 + lane], mine);
                }
                if (col_ok && cmax > 0) atomicMax(&P.colmax[col_c], cmax);
            } else if constexpr (EPI == EPI_F8_PROD) {
                // c = float2int_rn(acc) (exact small-integer sum), reduced once: r = c - p * mulhi(c, floor(2^32/p)) in (-p/2, 3p/2)
                const int pidx  = P.first_modulus + tc.unit / P.prods;
                const int32_t p = g8d_moduli[FP8][pidx], pinv = g8d_pinv32[FP8][pidx];
                int16_t *dst = reinterpret_cast<int16_t *>(P.out) + (size_t)tc.unit * P.out_stride + (size_t)col_c * P.ldc + row0;
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_COL; c0 += 32) {
                    int32_t v[32];
                    tmem_ld32(taddr0 + c0, v);
                    tmem_ld_wait();
                    uint32_t w[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int32_t c0i = __float2int_rn(__int_as_float(v[2 * j])), c1i = __float2int_rn(__int_as_float(v[2 * j + 1]));
                        const int32_t r0 = c0i - p * __mulhi(c0i, pinv), r1 = c1i - p * __mulhi(c1i, pinv);
                        w[j] = (uint32_t)(r0 & 0xFFFF) | ((uint32_t)r1 << 16);
                    }
                    if (col_ok) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4 *>(dst + c0 + 8 * j) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                    }
                }
            } else if constexpr (EPI == EPI_F8_MOD) {
                // mod.hpp:106-130: c_j = float2int_rn(acc_j); square moduli: sqrt(p)*(c0 + c1) + c2, else 256*c0 + 16*(c2 - c0 - c1) + c1; mod p
                const int32_t p = g8d_moduli[FP8][midx], pinv = g8d_pinv32[FP8][midx];
                const bool sq   = midx < 6;
                const int32_t sqrtp = sq ? g8d_f8sqrt[midx] : 0;
                int16_t *dst = reinterpret_cast<int16_t *>(P.out) + (size_t)tc.unit * P.out_stride + (size_t)col_c * P.ldc + row0;
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_COL; c0 += 16) {
                    int32_t a0[16], a1[16], a2[16];
                    tmem_ld16(taddr0 + c0, a0);
                    tmem_ld16(ta1 + c0, a1);
                    tmem_ld16(ta2 + c0, a2);
                    tmem_ld_wait();
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        int32_t o[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int32_t c0i = __float2int_rn(__int_as_float(a0[2 * j + e]));
                            const int32_t c1i = __float2int_rn(__int_as_float(a1[2 * j + e]));
                            const int32_t c2i = __float2int_rn(__int_as_float(a2[2 * j + e]));
                            const int32_t r0 = c0i - p * __mulhi(c0i, pinv), r1 = c1i - p * __mulhi(c1i, pinv), r2 = c2i - p * __mulhi(c2i, pinv);
                            const int32_t t  = sq ? sqrtp * (r0 + r1) + r2 : (r0 * 256) + ((r2 - r0 - r1) * 16) + r1;
                            o[e]             = mod_i32(t, p, pinv);
                        }
                        w[j] = (uint32_t)(o[0] & 0xFFFF) | ((uint32_t)o[1] << 16);
                    }
                    if (col_ok) {
                        *reinterpret_cast<uint4 *>(dst + c0)     = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4 *>(dst + c0 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                }
            } else if constexpr (EPI == EPI_F8_BOUND_CPLX) {
                // each accumulator summed 2k products in f32: inflate by (2k + 1) 2^-24 (the reference's bound chain, find_max.hpp:116-140,
                // differs in shape -- three GEMMs and round-up additions -- but FP8 shifts are not a bit-parity contract)
                const float ku2 = 2.0f * P.inflate;
                int32_t cmax = 0;
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_COL; c0 += 32) {
                    int32_t v0[32], v1[32];
                    tmem_ld32(taddr0 + c0, v0);
                    tmem_ld32(ta1 + c0, v1);
                    tmem_ld_wait();
                    int32_t mine = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float f0 = __int_as_float(v0[j]), f1 = __int_as_float(v1[j]);
                        const float u  = fmaxf(fmaxf(__fmaf_ru(ku2, f0, f0), __fmaf_ru(ku2, f1, f1)), 0.0f);
                        const int32_t x = __float_as_int(u);
                        cmax            = max(cmax, x);
                        const int32_t r = __reduce_max_sync(0xffffffffu, x);
                        mine            = (j == lane) ? r : mine;
                    }
                    if (mine > 0) atomicMax(&P.rowmax[row0 + c0 + lane], mine);
                }
                if (col_ok && cmax > 0) atomicMax(&P.colmax[col_c], cmax);
            } else { // EPI_BOUND_MAX_CPLX
                int32_t cmax = 0;
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_COL; c0 += 32) {
                    int32_t v0[32], v1[32];
                    tmem_ld32(taddr0 + c0, v0);
                    tmem_ld32(ta1 + c0, v1);
                    tmem_ld_wait();
                    int32_t mine = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int32_t x = max(v0[j], v1[j]);
                        cmax            = max(cmax, x);
                        const int32_t r = __reduce_max_sync(0xffffffffu, x);
                        mine            = (j == lane) ? r : mine;
                    }
                    if (mine > 0) atomicMax(&P.rowmax[row0 + c0 + lane], mine);
                }
                if (col_ok && cmax > 0) atomicMax(&P.colmax[col_c], cmax);
            }

            // all TMEM reads of this warp are complete (wait::ld above): hand the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
#pragma unroll
            for (int a = 0; a < EC::NACC; ++a) {
                if (lane == 0) {
                    if constexpr (CG == 2) mbar_arrive_cluster(mapa(smem_u32(&tempty_bar[buf]), 0)); // the MMA warp lives in the leader CTA
                    else mbar_arrive(&tempty_bar[buf]);
                }
                if (++buf == NUM_BUF) buf = 0, tphase ^= 1;
            }
        }
    }

    if constexpr (EPI == EPI_MOD_I8_SCATTER || EPI == EPI_RAW_I32_SCATTER) {
        if (warp >= 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // all bulk stores of this thread have been written
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); // neither CTA may exit (or free TMEM) while the pair's MMAs / remote arrives are in flight
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encoder() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

// 3-D view {k_pad bytes, rows, planes} of a stack of K-major int8 planes; box = {128, box_rows, 1}.
// `rows` is the VALID extent: rows beyond it are zero-filled by TMA, which is what makes the padded
// part of every tile contribute exact zeros (the reference leaves that padding uninitialised).
static bool make_plane_map(CUtensorMap *map, const void *base, size_t k_pad, size_t rows, size_t planes, size_t plane_stride,
                           int box_rows) {
    PFN_encodeTiled enc = get_encoder();
    if (!enc) return false;
    cuuint64_t dims[3]    = {(cuuint64_t)k_pad, (cuuint64_t)rows, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)k_pad, (cuuint64_t)plane_stride};
    cuuint32_t box[3]     = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3]    = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// store view of one owner's receive area: {ldc bytes of a column, owner_cols columns, units}, box = 128 rows x 32 columns (one epilogue
// warp's staging buffer), SWIZZLE_128B to match the conflict-free staging layout
static bool make_store_map(CUtensorMap *map, void *base, size_t ldc, size_t cols, size_t units, size_t unit_stride) {
    PFN_encodeTiled enc = get_encoder();
    if (!enc) return false;
    cuuint64_t dims[3]    = {(cuuint64_t)ldc, (cuuint64_t)cols, (cuuint64_t)units};
    cuuint64_t strides[2] = {(cuuint64_t)ldc, (cuuint64_t)(units > 1 ? unit_stride : ldc * cols)};
    cuuint32_t box[3]     = {128, 32, 1};
    cuuint32_t estr[3]    = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int num_sms() {
    static std::atomic<int> n[64]; // racing first calls store the same value
    int dev = 0;
    cudaGetDevice(&dev);
    int v = n[dev & 63].load(std::memory_order_relaxed);
    if (!v) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}

// G8_GEMM_CTA_GROUP=1 forces the single-CTA kernel (debugging / A-B comparisons); default is the CTA-pair kernel
static int cta_group_pref() {
    static int v = [] {
        const char *e = getenv("G8_GEMM_CTA_GROUP");
        return (e && e[0] == '1') ? 1 : 2;
    }();
    return v;
}

template <int EPI, int CG> static int launch_tc_cg(const GemmArgs &g, cudaStream_t st) {
    using KS = KernelShape<EPI, CG>;
    int planes = g.num_units;
    const int mods = ((EPI == EPI_MOD_I8 || EPI == EPI_MOD_I8_SCATTER) && g.prods > 1) ? (g.num_units + g.prods - 1) / g.prods : g.num_units;
    for (int i = 0; i < 3; ++i) planes = max(planes, max(g.groupA[i], g.groupB[i]) + mods);
    if (EPI == EPI_BOUND_MAX || EPI == EPI_F8_BOUND) planes = max(planes, max(g.groupA[0], g.groupB[0]) + g.num_units - 1 + max(1, g.kchain));
    if (EPI == EPI_BOUND_MAX_CPLX || EPI == EPI_F8_BOUND_CPLX) planes = max(planes, 2 * max(1, g.kchain));
    if (EPI == EPI_F8_MOD) planes = max(g.groupA[0], g.groupB[0]) + f8_plane_base(g.first_modulus + g.num_units);
    if (EPI == EPI_F8_PROD) planes = (g.prods / 3 - 1) * g.set_stride + f8_plane_base(g.first_modulus + (g.num_units + g.prods - 1) / g.prods);
    CUtensorMap mapL, mapC;
    if (!make_plane_map(&mapL, g.B, g.k_pad, g.n, planes, g.strideB, TILE_LANE)) return (int)cudaErrorNotSupported;
    if (!make_plane_map(&mapC, g.A, g.k_pad, g.m, planes, g.strideA, KS::TILE_COL / CG)) return (int)cudaErrorNotSupported;

    KParams P{};
    PeerMaps PM{};
    P.tiles_l       = (int)((g.n + TILE_LANE * CG - 1) / (TILE_LANE * CG));
    P.tiles_c       = (int)((g.m + KS::TILE_COL - 1) / KS::TILE_COL);
    P.num_units     = g.num_units;
    P.first_modulus = g.first_modulus;
    P.kblocks       = (int)(g.k_pad / BLOCK_K);
    P.n = (int)g.n, P.m = (int)g.m;
    for (int i = 0; i < 3; ++i) P.groupA[i] = g.groupA[i], P.groupB[i] = g.groupB[i];
    P.out = g.out, P.out_stride = g.out_stride, P.ldc = g.ldc;
    P.rowmax = g.rowmax, P.colmax = g.colmax;
    P.inflate = (float)(g.k_true + 1) * 0x1p-24f;
    P.owner_cols = 0, P.tl_rot = 0;
    P.kchain = g.kchain > 0 ? g.kchain : 1;
    static const int l2hint_pref = [] { const char *e = getenv("G8_GEMM_L2HINT"); return e ? atoi(e) : 4; }(); // default: streaming C_mid stores (r02h)
    P.l2hint = l2hint_pref;
    static const int group_pref = [] { const char *e = getenv("G8_GEMM_GROUP"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 16; }();
    P.group = group_pref;
    P.prods = g.prods > 0 ? g.prods : (EPI == EPI_F8_PROD ? 3 : 1), P.set_stride = g.set_stride;
    if (g.owner_cols) {
        // the scatter is tile-granular: every lane tile (TILE_LANE * CG columns) must fall inside one owner's slab
        if (g.owner_cols % (TILE_LANE * CG) || (EPI != EPI_MOD_I8_SCATTER && EPI != EPI_RAW_I32_SCATTER)) return (int)cudaErrorInvalidValue;
        P.owner_cols = (int)g.owner_cols;
        for (int i = 0; i < G8_MAX_PEERS; ++i) P.peer_out[i] = g.peer_out[i];
        if (EPI == EPI_MOD_I8_SCATTER) {
            if (g.world < 1 || g.world > G8_MAX_PEERS || g.ldc % 16 || g.out_stride % 16) return (int)cudaErrorInvalidValue;
            for (int o = 0; o < g.world; ++o) {
                if (!g.peer_out[o] || (reinterpret_cast<uintptr_t>(g.peer_out[o]) & 15)) return (int)cudaErrorInvalidValue;
                if (!make_store_map(&PM.m[o], g.peer_out[o], g.ldc, g.owner_cols, (size_t)g.num_units, g.out_stride)) return (int)cudaErrorNotSupported;
            }
        }
        P.tl_rot = (int)(((size_t)(g.rank + 1) % (size_t)g.world) * (g.owner_cols / (TILE_LANE * CG))) % P.tiles_l;
    }

    // per kernel instantiation AND per device (function attributes are per-device state; one process may drive several GPUs)
    static std::atomic<bool> attr_set_dev[64];
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (!attr_set_dev[cur_dev & 63].load(std::memory_order_acquire)) { // setting it twice (two racing first calls) is harmless
        cudaError_t e = cudaFuncSetAttribute(gemm_i8_tc_kernel<EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, KS::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set_dev[cur_dev & 63].store(true, std::memory_order_release);
    }
    const int total = P.num_units * P.tiles_l * P.tiles_c;
    const int grid  = CG * min(total, num_sms() / CG);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(NUM_THREADS), cfg.dynamicSmemBytes = KS::SMEM_BYTES, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, gemm_i8_tc_kernel<EPI, CG>, mapL, mapC, P, PM);
}

template <int EPI> static int launch_tc(const GemmArgs &g, cudaStream_t st) {
    if (g.m == 0 || g.n == 0 || g.num_units == 0) return 0;
    return cta_group_pref() == 2 ? launch_tc_cg<EPI, 2>(g, st) : launch_tc_cg<EPI, 1>(g, st);
}

int launch_gemm_tc(const GemmArgs &g, cudaStream_t st) {
    switch (g.epi) {
    case EPI_MOD_I8: return g.owner_cols ? launch_tc<EPI_MOD_I8_SCATTER>(g, st) : launch_tc<EPI_MOD_I8>(g, st);
    case EPI_RAW_I32: return g.owner_cols ? launch_tc<EPI_RAW_I32_SCATTER>(g, st) : launch_tc<EPI_RAW_I32>(g, st);
    case EPI_BOUND_MAX: return launch_tc<EPI_BOUND_MAX>(g, st);
    case EPI_MOD_I8_CPLX: return launch_tc<EPI_MOD_I8_CPLX>(g, st);
    case EPI_BOUND_MAX_CPLX: return launch_tc<EPI_BOUND_MAX_CPLX>(g, st);
    case EPI_F8_MOD: return launch_tc<EPI_F8_MOD>(g, st);
    case EPI_F8_BOUND: return launch_tc<EPI_F8_BOUND>(g, st);
    case EPI_F8_RAW: return launch_tc<EPI_F8_RAW>(g, st);
    case EPI_F8_BOUND_CPLX: return launch_tc<EPI_F8_BOUND_CPLX>(g, st);
    case EPI_F8_PROD: return launch_tc<EPI_F8_PROD>(g, st);
    }
    return (int)cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------
// TEST/DEBUG ONLY: straightforward dp4a kernel with the same epilogues, used by tests to cross-check the
// tensor-core kernel on the GPU.  g8_gemm() never calls it.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t dot_k(const int8_t *a, const int8_t *b, size_t k_pad) {
    const int4 *pa = reinterpret_cast<const int4 *>(a), *pb = reinterpret_cast<const int4 *>(b);
    int32_t acc = 0;
    for (size_t i = 0; i < k_pad / 16; ++i) {
        const int4 x = pa[i], y = pb[i];
        acc = __dp4a(x.x, y.x, acc);
        acc = __dp4a(x.y, y.y, acc);
        acc = __dp4a(x.z, y.z, acc);
        acc = __dp4a(x.w, y.w, acc);
    }
    return acc;
}

__global__ void gemm_i8_simt_kernel(GemmArgs g) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; // row of C
    const size_t c = blockIdx.y;                                    // column of C
    const int u    = blockIdx.z;
    if (r >= g.m || c >= g.n) return;
    auto A = [&](int grp) { return g.A + (size_t)(g.groupA[grp] + u) * g.strideA + r * g.k_pad; };
    auto B = [&](int grp) { return g.B + (size_t)(g.groupB[grp] + u) * g.strideB + c * g.k_pad; };
    const int midx  = g.first_modulus + u;
    const int32_t p = g8d_moduli[INT8][midx], pinv = g8d_pinv32[INT8][midx];
    const size_t o  = (size_t)u * g.out_stride + c * g.ldc + r;
    switch (g.epi) {
    case EPI_MOD_I8: reinterpret_cast<int8_t *>(g.out)[o] = (int8_t)mod_i32(dot_k(A(0), B(0), g.k_pad), p, pinv); break;
    case EPI_RAW_I32: reinterpret_cast<int32_t *>(g.out)[o] = dot_k(A(0), B(0), g.k_pad); break;
    case EPI_BOUND_MAX: {
        const int32_t v = dot_k(A(0), B(0), g.k_pad);
        atomicMax(&g.rowmax[r], v);
        atomicMax(&g.colmax[c], v);
    } break;
    case EPI_MOD_I8_CPLX: {
        const int64_t rr = dot_k(A(0), B(0), g.k_pad), ii = dot_k(A(1), B(1), g.k_pad), ri = dot_k(A(2), B(2), g.k_pad);
        const int64_t re = rr - ii, im = ri - rr - ii;
        int8_t *out = reinterpret_cast<int8_t *>(g.out) + 2 * o;
        out[0] = (int8_t)mod_i64(re, p, g8d_pinv64[INT8][midx]);
        out[1] = (int8_t)mod_i64(im, p, g8d_pinv64[INT8][midx]);
    } break;
    case EPI_BOUND_MAX_CPLX: {
        const int32_t v0 = dot_k(A(0), B(0), g.k_pad) + dot_k(A(1), B(1), g.k_pad);
        const int32_t v1 = dot_k(A(0), B(1), g.k_pad) + dot_k(A(1), B(0), g.k_pad);
        const int32_t v  = max(v0, v1);
        atomicMax(&g.rowmax[r], v);
        atomicMax(&g.colmax[c], v);
    } break;
    }
}

int launch_gemm_simt(const GemmArgs &g, cudaStream_t st) {
    if (g.m == 0 || g.n == 0 || g.num_units == 0) return 0;
    const dim3 grid((unsigned)((g.m + 127) / 128), (unsigned)g.n, (unsigned)g.num_units);
    gemm_i8_simt_kernel<<<grid, 128, 0, st>>>(g);
    return (int)cudaGetLastError();
}

} // namespace g8
