// gemmul8_b200 -- synthetic test matrices, same generator as the reference harness
// (testing/make_matrix.hpp:33-82): element idx draws from curand_init(seed, idx, 0): first a uniform
// double u, then a normal double g (complex: u_r, u_i, g_r, g_i); phi < 0 -> g, else (u - 0.5) * exp(g * phi).
// Measurement support only (bench.py, tests); not part of the GEMM hot path.
#include "g8_internal.cuh"
#include "../../include/gemmul8_c.h"
#include <curand_kernel.h>

namespace g8 {
template <typename T> __global__ void randmat_kernel(T *X, size_t count, double phi, unsigned long long seed) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    curandState state;
    curand_init(seed, idx, 0, &state);
    if constexpr (sizeof(T) == 2 * sizeof(decltype(T{}.x)) && !std::is_arithmetic<T>::value) {
        using U         = decltype(T{}.x);
        const double ur = curand_uniform_double(&state), ui = curand_uniform_double(&state);
        const double gr = curand_normal_double(&state), gi = curand_normal_double(&state);
        T out;
        if (phi < 0) out.x = (U)gr, out.y = (U)gi;
        else out.x = (U)((ur - 0.5) * exp(gr * phi)), out.y = (U)((ui - 0.5) * exp(gi * phi));
        X[idx] = out;
    }
}
template <typename T> __global__ void randmat_real_kernel(T *X, size_t count, double phi, unsigned long long seed) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    curandState state;
    curand_init(seed, idx, 0, &state);
    const double u = curand_uniform_double(&state);
    const double g = curand_normal_double(&state);
    X[idx]         = (phi < 0) ? (T)g : (T)((u - 0.5) * exp(g * phi));
}
} // namespace g8

extern "C" __attribute__((visibility("default"))) int g8_randmat(int dtype, void *X, size_t rows, size_t cols, double phi, unsigned long long seed, void *stream) {
    const size_t count = rows * cols;
    if (!X || dtype < 0 || dtype > 3) return G8_STATUS_INVALID_VALUE;
    if (count == 0) return 0;
    cudaStream_t st     = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((count + 255) / 256);
    switch (dtype) {
    case g8::F32: g8::randmat_real_kernel<float><<<grid, 256, 0, st>>>(static_cast<float *>(X), count, phi, seed); break;
    case g8::F64: g8::randmat_real_kernel<double><<<grid, 256, 0, st>>>(static_cast<double *>(X), count, phi, seed); break;
    case g8::C32: g8::randmat_kernel<float2><<<grid, 256, 0, st>>>(static_cast<float2 *>(X), count, phi, seed); break;
    default: g8::randmat_kernel<double2><<<grid, 256, 0, st>>>(static_cast<double2 *>(X), count, phi, seed); break;
    }
    return (int)cudaGetLastError();
}
