// gemmul8_b200 -- stage 3: CRT accumulation, modular fold, unscale and alpha/beta epilogue in one HBM pass.
//
// Replaces reference inverse_scaling_real.hpp:8-278 and inverse_scaling_complex.hpp:8-326.
// The floating-point operation ORDER is contractual for bit parity and is kept exactly:
//   single chain (T float, or N <= P_is_double):  S = fma(w_i, c_i, S), i ascending; q = rint(invP*S);
//                                                 r = fma(P.x, q, S)
//   hi/lo chain (T double, N > P_is_double):      H = fma(w_i.x, c_i, H); L = fma(w_i.y, c_i, L);
//                                                 q = rint(invP*H); r = fma(P.y, q, fma(P.x, q, H) + L)
//   AB = scalbn((T)r, sftA[row] + sftB[col]);  then C = AB | C+AB | -AB | C-AB | fma(beta, C, alpha*AB).
// What changes is the memory access and the instruction mix: each thread owns 8 consecutive residues of one column
// (one 64-bit load per plane, 128-bit stores; the per-plane address / constant / loop overhead is shared by 8 results),
// and int8 -> double goes through a mantissa splice + one DADD instead of I2F.F64 (quarter-rate pipe), instead of one
// element per thread with N strided byte loads.
#include "g8_internal.cuh"

namespace g8 {

template <typename T> struct CrtTraits;
template <> struct CrtTraits<float>   { using U = float;  static constexpr bool cplx = false; };
template <> struct CrtTraits<double>  { using U = double; static constexpr bool cplx = false; };
template <> struct CrtTraits<float2>  { using U = float;  static constexpr bool cplx = true;  };
template <> struct CrtTraits<double2> { using U = double; static constexpr bool cplx = true;  };

// v * 2^s.  The reference calls scalbn()/scalbnf(); CUDA's scalbn is a single multiplication by the exactly
// representable 2^s whenever |s| <= 1021, which we issue directly (bit-identical); otherwise defer to the library.
__device__ __forceinline__ double scal(double v, int s) {
    if (abs(s) <= 1021) return __dmul_rn(v, __longlong_as_double((long long)(1023 + s) << 52));
    return scalbn(v, s);
}
__device__ __forceinline__ float  scal(float v, int s) { return scalbnf(v, s); }
__device__ __forceinline__ float  fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }

// int8 / int16 residue (already XORed with the sign bit) -> double without a conversion instruction and without re-materialising the
// constant high word: `tmpl` is a register pair whose high word stays 0x43300000 (2^52); PRMT rewrites only its low word with the
// selected byte(s), and ONE DADD removes 2^52 + bias.  (Written as asm because the compiler otherwise emits PRMT + MOV per value.)
template <int SEL> __device__ __forceinline__ double splice_low(double &tmpl, uint32_t w, double off) {
    asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tprmt.b32 lo, %1, 0, %2;\n\tmov.b64 %0, {lo, hi};\n\t}" : "+d"(tmpl) : "r"(w), "n"(SEL));
    return __dadd_rn(tmpl, off);
}

struct Scalars {
    double ar, ai, br, bi; // host scalars widened (exact); used in MODE 4
};

constexpr int CRT_NV = 8; // residue bytes (= scalar outputs) per thread: one 64-bit load per plane (256 B per warp)

// MODE: 0 C=AB, 1 C+=AB, 2 C=-AB, 3 C-=AB, 4 general (host scalars), 5 general (device scalars)
// PARTS: 0 = plain C_mid; 2 / 4 / 8 = K-sharded, up to that many per-shard residue arrays (c.nparts of them are real)
template <typename T, bool DD, int MODE, bool VECIO, int BE, int PARTS = 0>
__global__ void __launch_bounds__(256) crt_kernel(CrtArgs c, Scalars hs, size_t groups, size_t total) {
    using TR           = CrtTraits<T>;
    using U            = typename TR::U;
    constexpr int NV   = CRT_NV;
    constexpr int VEC  = TR::cplx ? NV / 2 : NV; // rows per thread
    const size_t idx   = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    // (32-bit division whenever the index fits: the emulated 64-bit one costs ~60 instructions per thread)
    const size_t col = (total <= 0xFFFFFFFFull) ? (size_t)((uint32_t)idx / (uint32_t)groups) : idx / groups;
    const size_t row = (idx - col * groups) * VEC;
    const int N      = c.num_moduli;

    // ---- accumulate over moduli (i ascending).  int8 -> double without the slow I2F path: the byte c+128 is
    //      spliced into the mantissa of 2^52 and (2^52 + 128) is subtracted (exact). ----
    double hi[NV], lo[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) hi[j] = 0.0, lo[j] = 0.0;
    constexpr int MIDB   = (BE == INT8) ? 1 : 2; // bytes per residue: int8 (INT8 backend) / int16 (FP8 backend)
    const int8_t *src = reinterpret_cast<const int8_t *>(c.C_mid) + (col * c.ldmid + row) * (TR::cplx ? 2 : 1) * MIDB;
    const size_t pstride = c.plane_stride * (TR::cplx ? 2 : 1) * MIDB;
    const int tbl1 = N - 2, tbl2 = N - thresholds(BE).P_is_double - 1;
    constexpr double kOff = 4503599627370496.0 + (BE == INT8 ? 128.0 : 32768.0);
    // (the high word is made opaque to the compiler -- total >> 63 is 0 -- so that it is kept in the register pair instead of being
    //  re-materialised with a MOV in front of every conversion)
    const int hi52 = 0x43300000 | (int)(total >> 63);
    double tmplA = __hiloint2double(hi52, 0), tmplB = __hiloint2double(hi52, 1); // see splice_low
    const int8_t *sp = src; // plane i: sp = src + i * pstride, advanced by addition (no 64-bit multiply per plane)
    // planes in flight per thread: 4 for the plain C_mid; K-sharded: PARTS loads per plane, so fewer planes (>= 8 loads in flight)
    constexpr int UNR = PARTS >= 8 ? 1 : PARTS >= 4 ? 2 : 4;
#pragma unroll UNR
    for (int i = 0; i < N; ++i, sp += pstride) {
        double cd[NV];
        if constexpr (BE == INT8) {
            uint32_t w[NV / 4];
            if constexpr (PARTS == 0) {
                const uint2 v = __ldg(reinterpret_cast<const uint2 *>(sp));
                w[0] = v.x, w[1] = v.y;
            } else {
                // K-sharded: add the per-shard residues byte-wise (dp4a against one-hot selectors sign-extends and adds in one
                // instruction), reduce mod p_i again -- what g8_stage_residue_sum does, without the round trip through HBM.
                // Constant trip count (PARTS, predicated) so that the loads of all shards are in flight together.
                uint2 x[PARTS > 0 ? PARTS : 1];
#pragma unroll
                for (int q = 0; q < PARTS; ++q)
                    x[q] = (q < c.nparts) ? __ldcs(reinterpret_cast<const uint2 *>(sp + (size_t)q * c.part_stride)) : make_uint2(0u, 0u);
                const int32_t p = g8d_moduli[INT8][i], pinv = g8d_pinv32[INT8][i];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                    for (int q = 0; q < PARTS; ++q) {
                        const int xv = (int)(h ? x[q].y : x[q].x);
                        a0 = __dp4a(xv, 0x00000001, a0), a1 = __dp4a(xv, 0x00000100, a1);
                        a2 = __dp4a(xv, 0x00010000, a2), a3 = __dp4a(xv, 0x01000000, a3);
                    }
                    a0 = mod_i32(a0, p, pinv), a1 = mod_i32(a1, p, pinv), a2 = mod_i32(a2, p, pinv), a3 = mod_i32(a3, p, pinv);
                    w[h] = (uint32_t)(a0 & 0xFF) | ((uint32_t)(a1 & 0xFF) << 8) | ((uint32_t)(a2 & 0xFF) << 16) | ((uint32_t)a3 << 24);
                }
            }
            w[0] ^= 0x80808080u, w[1] ^= 0x80808080u;
            cd[0] = splice_low<0x4440>(tmplA, w[0], -kOff), cd[1] = splice_low<0x4441>(tmplB, w[0], -kOff);
            cd[2] = splice_low<0x4442>(tmplA, w[0], -kOff), cd[3] = splice_low<0x4443>(tmplB, w[0], -kOff);
            cd[4] = splice_low<0x4440>(tmplA, w[1], -kOff), cd[5] = splice_low<0x4441>(tmplB, w[1], -kOff);
            cd[6] = splice_low<0x4442>(tmplA, w[1], -kOff), cd[7] = splice_low<0x4443>(tmplB, w[1], -kOff);
        } else {
            uint4 w = __ldg(reinterpret_cast<const uint4 *>(sp));
            const uint32_t ww[4] = {w.x ^ 0x80008000u, w.y ^ 0x80008000u, w.z ^ 0x80008000u, w.w ^ 0x80008000u};
            cd[0] = splice_low<0x4410>(tmplA, ww[0], -kOff), cd[1] = splice_low<0x4432>(tmplB, ww[0], -kOff);
            cd[2] = splice_low<0x4410>(tmplA, ww[1], -kOff), cd[3] = splice_low<0x4432>(tmplB, ww[1], -kOff);
            cd[4] = splice_low<0x4410>(tmplA, ww[2], -kOff), cd[5] = splice_low<0x4432>(tmplB, ww[2], -kOff);
            cd[6] = splice_low<0x4410>(tmplA, ww[3], -kOff), cd[7] = splice_low<0x4432>(tmplB, ww[3], -kOff);
        }
        if constexpr (DD) {
            const double wx = g8d_qPi2[BE][tbl2][i][0], wy = g8d_qPi2[BE][tbl2][i][1];
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                hi[j] = fma(wx, cd[j], hi[j]);
                lo[j] = fma(wy, cd[j], lo[j]);
            }
        } else {
            const double wv = g8d_qPi1[BE][tbl1][i];
#pragma unroll
            for (int j = 0; j < NV; ++j) hi[j] = fma(wv, cd[j], hi[j]);
        }
    }

    // ---- fold modulo P, cast, unscale ----
    const double invP = g8d_invP[BE][N - 2];
    const double Px = g8d_P[BE][N - 2][0], Py = g8d_P[BE][N - 2][1];
    const int sB = c.sftB[col];
    // sftA has pad256(m) entries, so the (unused) tail rows of the last group may be read safely; row % VEC == 0
    __align__(16) int16_t sA[VEC];
    if constexpr (VEC == 8) *reinterpret_cast<uint4 *>(sA) = *reinterpret_cast<const uint4 *>(c.sftA + row);
    else *reinterpret_cast<uint2 *>(sA) = *reinterpret_cast<const uint2 *>(c.sftA + row);
    U ab[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int sft   = (int)sA[TR::cplx ? j / 2 : j] + sB;
        const double q  = rint(invP * hi[j]);
        double r;
        if constexpr (DD) r = fma(Py, q, fma(Px, q, hi[j]) + lo[j]);
        else r = fma(Px, q, hi[j]);
        ab[j] = scal((U)r, sft);
    }

    // ---- alpha / beta ----
    U *dst = reinterpret_cast<U *>(c.C) + (col * c.ldc + row) * (TR::cplx ? 2 : 1);
    const int valid = (int)min((size_t)VEC, c.m - row) * (TR::cplx ? 2 : 1);
    constexpr int PER16 = 16 / sizeof(U); // scalars per 128-bit access
    const bool full = VECIO && valid == NV;
    // general modes: the scalars first -- beta == 0 means "C is not read" (the BLAS convention cuBLAS follows; the reference's
    // fma(beta, C, alpha * AB) would turn a stale NaN / Inf in an uninitialised C into a NaN result, inverse_scaling_real.hpp:171-186)
    U ar = 1, ai = 0, br = 0, bi = 0;
    bool read_old = (MODE == 1 || MODE == 3);
    if constexpr (MODE >= 4) {
        if constexpr (MODE == 5) {
            const U *pa = reinterpret_cast<const U *>(c.alpha), *pb = reinterpret_cast<const U *>(c.beta);
            ar = pa[0], br = pb[0];
            if constexpr (TR::cplx) ai = pa[1], bi = pb[1];
        } else {
            ar = (U)hs.ar, ai = (U)hs.ai, br = (U)hs.br, bi = (U)hs.bi;
        }
        read_old = !(br == U(0) && bi == U(0));
    }
    __align__(16) U old[NV];
    if constexpr (MODE == 1 || MODE == 3 || MODE >= 4) {
        if (!read_old) {
#pragma unroll
            for (int j = 0; j < NV; ++j) old[j] = U(0);
        } else if (full) {
#pragma unroll
            for (int j = 0; j < NV; j += PER16) *reinterpret_cast<uint4 *>(old + j) = *reinterpret_cast<const uint4 *>(dst + j);
        } else {
#pragma unroll
            for (int j = 0; j < NV; ++j) old[j] = (j < valid) ? dst[j] : U(0);
        }
    }
    __align__(16) U out[NV];
    if constexpr (MODE == 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j) out[j] = ab[j];
    } else if constexpr (MODE == 1) {
#pragma unroll
        for (int j = 0; j < NV; ++j) out[j] = old[j] + ab[j];
    } else if constexpr (MODE == 2) {
#pragma unroll
        for (int j = 0; j < NV; ++j) out[j] = -ab[j];
    } else if constexpr (MODE == 3) {
#pragma unroll
        for (int j = 0; j < NV; ++j) out[j] = old[j] - ab[j];
    } else {
        if constexpr (TR::cplx) {
            // Taxpby_scal (template_math.hpp:64-75)
#pragma unroll
            for (int j = 0; j < NV; j += 2) {
                const U xr = ab[j], xi = ab[j + 1], yr = old[j], yi = old[j + 1];
                out[j]     = fma_(-bi, yi, fma_(br, yr, fma_(-ai, xi, ar * xr)));
                out[j + 1] = fma_(bi, yr, fma_(br, yi, fma_(ai, xr, ar * xi)));
            }
        } else {
#pragma unroll
            for (int j = 0; j < NV; ++j) out[j] = fma_(br, old[j], ar * ab[j]);
        }
    }
    if (full) {
#pragma unroll
        for (int j = 0; j < NV; j += PER16) *reinterpret_cast<uint4 *>(dst + j) = *reinterpret_cast<const uint4 *>(out + j);
    } else {
#pragma unroll
        for (int j = 0; j < NV; ++j)
            if (j < valid) dst[j] = out[j];
    }
}

template <typename T, bool DD, int MODE> static void crt_go(const CrtArgs &c, const Scalars &hs, cudaStream_t st) {
    constexpr int VEC   = CrtTraits<T>::cplx ? CRT_NV / 2 : CRT_NV;
    const size_t groups = (c.m + VEC - 1) / VEC;
    const size_t total  = groups * c.n;
    if (total == 0) return;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(c.C) % 16 == 0) && ((c.ldc * sizeof(T)) % 16 == 0);
    const unsigned grid = (unsigned)((total + 255) / 256);
    if constexpr (!CrtTraits<T>::cplx) {
        if (c.nparts > 1) { // K-sharded: per-shard residues summed inside the kernel (real types, INT8 backend; checked by the C ABI)
            auto go = [&](auto kv, auto kn) { vec_ok ? kv<<<grid, 256, 0, st>>>(c, hs, groups, total) : kn<<<grid, 256, 0, st>>>(c, hs, groups, total); };
            if (c.nparts <= 2) go(crt_kernel<T, DD, MODE, true, INT8, 2>, crt_kernel<T, DD, MODE, false, INT8, 2>);
            else if (c.nparts <= 4) go(crt_kernel<T, DD, MODE, true, INT8, 4>, crt_kernel<T, DD, MODE, false, INT8, 4>);
            else go(crt_kernel<T, DD, MODE, true, INT8, 8>, crt_kernel<T, DD, MODE, false, INT8, 8>);
            return;
        }
    }
    if (c.backend == FP8) {
        if (vec_ok) crt_kernel<T, DD, MODE, true, FP8><<<grid, 256, 0, st>>>(c, hs, groups, total);
        else crt_kernel<T, DD, MODE, false, FP8><<<grid, 256, 0, st>>>(c, hs, groups, total);
    } else {
        if (vec_ok) crt_kernel<T, DD, MODE, true, INT8><<<grid, 256, 0, st>>>(c, hs, groups, total);
        else crt_kernel<T, DD, MODE, false, INT8><<<grid, 256, 0, st>>>(c, hs, groups, total);
    }
}

template <typename T, bool DD> static void crt_mode(const CrtArgs &c, int mode, const Scalars &hs, cudaStream_t st) {
    switch (mode) {
    case 0: crt_go<T, DD, 0>(c, hs, st); break;
    case 1: crt_go<T, DD, 1>(c, hs, st); break;
    case 2: crt_go<T, DD, 2>(c, hs, st); break;
    case 3: crt_go<T, DD, 3>(c, hs, st); break;
    case 4: crt_go<T, DD, 4>(c, hs, st); break;
    default: crt_go<T, DD, 5>(c, hs, st); break;
    }
}

// alpha/beta residency and special cases follow inverse_scaling_real.hpp:209-237 / _complex.hpp:254-284:
// device-resident alpha -> general kernel reading the scalars on the device; host scalars -> the four
// special cases (alpha = +-1, beta in {0,1}) else the general FMA form.
int launch_crt(const CrtArgs &c, int dtype, cudaStream_t st) {
    cudaPointerAttributes attr{};
    cudaError_t e = cudaPointerGetAttributes(&attr, c.alpha);
    if (e != cudaSuccess) cudaGetLastError(); // plain host pointers may report an error on old drivers: treat as host
    const bool is_device = (e == cudaSuccess) && attr.type != cudaMemoryTypeUnregistered && attr.type != cudaMemoryTypeHost;
    Scalars hs{1, 0, 0, 0};
    int mode = 5;
    if (!is_device) {
        switch (dtype) {
        case F32: hs.ar = *static_cast<const float *>(c.alpha); hs.br = *static_cast<const float *>(c.beta); break;
        case F64: hs.ar = *static_cast<const double *>(c.alpha); hs.br = *static_cast<const double *>(c.beta); break;
        case C32: hs.ar = static_cast<const float *>(c.alpha)[0]; hs.ai = static_cast<const float *>(c.alpha)[1];
                  hs.br = static_cast<const float *>(c.beta)[0];  hs.bi = static_cast<const float *>(c.beta)[1]; break;
        default:  hs.ar = static_cast<const double *>(c.alpha)[0]; hs.ai = static_cast<const double *>(c.alpha)[1];
                  hs.br = static_cast<const double *>(c.beta)[0];  hs.bi = static_cast<const double *>(c.beta)[1]; break;
        }
        mode = 4;
        if (hs.ai == 0.0 && hs.bi == 0.0) {
            if (hs.ar == 1.0 && hs.br == 0.0) mode = 0;
            else if (hs.ar == 1.0 && hs.br == 1.0) mode = 1;
            else if (hs.ar == -1.0 && hs.br == 0.0) mode = 2;
            else if (hs.ar == -1.0 && hs.br == 1.0) mode = 3;
        }
    }
    const bool dd = c.num_moduli > thresholds(c.backend).P_is_double;
    switch (dtype) {
    case F32: crt_mode<float, false>(c, mode, hs, st); break;
    case C32: crt_mode<float2, false>(c, mode, hs, st); break;
    case F64: dd ? crt_mode<double, true>(c, mode, hs, st) : crt_mode<double, false>(c, mode, hs, st); break;
    default:  dd ? crt_mode<double2, true>(c, mode, hs, st) : crt_mode<double2, false>(c, mode, hs, st); break;
    }
    return (int)cudaGetLastError();
}

} // namespace g8
