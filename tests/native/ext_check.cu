// Drop-in check of include/gemmul8_ext.hpp: the C++ front ends of the host-buffer pipeline (HostGemm) and of the K-sharded multi-GPU
// driver (MgComm / MgGemm, here with a world of one rank) must reproduce gemmul8::gemmLt bit for bit in accurate mode.
// Links against lib/libgemmul8.a exactly like a user of the reference's static library would.
#include "../../include/gemmul8_ext.hpp"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#define CK(x)                                                                    \
    do {                                                                         \
        const int _e = (int)(x);                                                 \
        if (_e != 0) {                                                           \
            std::fprintf(stderr, "%s -> %d (line %d)\n", #x, _e, __LINE__);      \
            return 1;                                                            \
        }                                                                        \
    } while (0)

static uint64_t lcg(uint64_t &s) { return s = s * 6364136223846793005ull + 1442695040888963407ull; }
static double unit(uint64_t &s) { return (double)(int64_t)(lcg(s) >> 11) * 0x1p-52 - 1.0; } // (-1, 1)
template <typename T> static T make(uint64_t &s);
template <> double make<double>(uint64_t &s) { return unit(s) * (1.0 + 1000.0 * unit(s) * unit(s)); }
template <> float make<float>(uint64_t &s) { return (float)unit(s); }
template <> cuFloatComplex make<cuFloatComplex>(uint64_t &s) { return make_cuFloatComplex((float)unit(s), (float)unit(s)); }
template <> cuDoubleComplex make<cuDoubleComplex>(uint64_t &s) { return make_cuDoubleComplex(unit(s), unit(s)); }
template <typename T> static T scalar(double re);
template <> double scalar<double>(double re) { return re; }
template <> float scalar<float>(double re) { return (float)re; }
template <> cuFloatComplex scalar<cuFloatComplex>(double re) { return make_cuFloatComplex((float)re, 0.0f); }
template <> cuDoubleComplex scalar<cuDoubleComplex>(double re) { return make_cuDoubleComplex(re, 0.0); }
template <typename T> struct is_cplx { static constexpr bool value = false; };
template <> struct is_cplx<cuFloatComplex> { static constexpr bool value = true; };
template <> struct is_cplx<cuDoubleComplex> { static constexpr bool value = true; };

template <typename T, gemmul8::Backend BE> static int run_case(const char *name, unsigned N, bool fast, cublasOperation_t opA, cublasOperation_t opB) {
    using namespace gemmul8;
    const size_t m = 300, n = 512, k = 640;
    const size_t rA = opA == CUBLAS_OP_N ? m : k, cA = opA == CUBLAS_OP_N ? k : m, rB = opB == CUBLAS_OP_N ? k : n, cB = opB == CUBLAS_OP_N ? n : k;
    uint64_t seed = 0x9E3779B97F4A7C15ull;
    std::vector<T> hA(rA * cA), hB(rB * cB), hC(m * n), want(m * n), got(m * n);
    for (auto &v : hA) v = make<T>(seed);
    for (auto &v : hB) v = make<T>(seed);
    T *dA, *dB, *dC;
    CK(cudaMalloc(&dA, sizeof(T) * hA.size()));
    CK(cudaMalloc(&dB, sizeof(T) * hB.size()));
    CK(cudaMalloc(&dC, sizeof(T) * m * n));
    CK(cudaMemcpy(dA, hA.data(), sizeof(T) * hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), sizeof(T) * hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0, sizeof(T) * m * n));
    const T one = scalar<T>(1.0), zero = scalar<T>(0.0);
    constexpr bool cplx = is_cplx<T>::value;

    // 1. the reference-compatible call
    void *work;
    CK(cudaMalloc(&work, workSize<cplx, BE>(m, n, k, N)));
    gemmLt<T, BE>(nullptr, opA, opB, m, n, k, &one, dA, rA, dB, rB, &zero, dC, m, N, fast, work);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(want.data(), dC, sizeof(T) * m * n, cudaMemcpyDeviceToHost));

    // 2. host buffers in / out
    ext::HostGemm<T, BE> host(opA, opB, m, n, k, N, fast, 256);
    CK(host.status());
    CK(host(&one, hA.data(), rA, hB.data(), rB, &zero, hC.data(), m));
    CK(cudaDeviceSynchronize());
    const bool same_host = std::memcmp(hC.data(), want.data(), sizeof(T) * m * n) == 0;

    // 3. the multi-GPU driver with a world of one rank (2 .. 8 ranks: mg_check.cu)
    bool same_mg = true;
    {
        ext::MgComm comm(1, 0, 8 * (m + n) + 4096);
        CK(comm.status());
        CK(comm.connect(comm.handle()));
        ext::MgGemm<T, BE> mg(comm, opA, opB, m, n, k, N, fast);
        CK(mg.status());
        CK(cudaMemset(dC, 0, sizeof(T) * m * n));
        CK(mg(&one, dA, rA, dB, rB, &zero, dC, m));
        CK(cudaDeviceSynchronize());
        CK(comm.health());
        CK(cudaMemcpy(got.data(), dC, sizeof(T) * m * n, cudaMemcpyDeviceToHost));
        same_mg = fast || std::memcmp(got.data(), want.data(), sizeof(T) * m * n) == 0; // fast mode: different statistics kernels (tolerance tested elsewhere)
    }
    std::printf("%s N=%u %s: HostGemm %s, MgGemm %s\n", name, N, fast ? "fast" : "accu", same_host ? "bit-identical" : "MISMATCH",
                fast ? "ran" : (same_mg ? "bit-identical" : "MISMATCH"));
    cudaFree(dA), cudaFree(dB), cudaFree(dC), cudaFree(work);
    return (same_host && same_mg) ? 0 : 1;
}

int main() {
    using gemmul8::Backend;
    int bad = 0;
    for (int fast = 0; fast < 2; ++fast) {
        bad += run_case<double, Backend::INT8>("DGEMM INT8", 14, fast, CUBLAS_OP_N, CUBLAS_OP_N);
        bad += run_case<double, Backend::FP8>("DGEMM FP8", 12, fast, CUBLAS_OP_T, CUBLAS_OP_N);
        bad += run_case<float, Backend::INT8>("SGEMM INT8", 6, fast, CUBLAS_OP_N, CUBLAS_OP_T);
        bad += run_case<cuFloatComplex, Backend::INT8>("CGEMM INT8", 7, fast, CUBLAS_OP_C, CUBLAS_OP_N);
        bad += run_case<cuDoubleComplex, Backend::FP8>("ZGEMM FP8", 9, fast, CUBLAS_OP_N, CUBLAS_OP_C);
    }
    std::printf(bad ? "ext_check FAILED (%d cases)\n" : "ext_check OK\n", bad);
    return bad ? 1 : 0;
}
