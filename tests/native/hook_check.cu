// Restated from the reference's debug/test_hijack.cu: plain cuBLAS calls; run once natively and once under
// LD_PRELOAD=libgemmul8.so with GEMMUL8_* set; prints a checksum line per call so the harness can compare both runs.
// With GEMMUL8_SKIP_SCALE_{A,B}=1 the same A/B are reused across calls and then replaced, to exercise cache invalidation
// (debug/test_hijack.cu:164-177).
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
static double checksum(const std::vector<double> &v) {
    double s = 0;
    for (size_t i = 0; i < v.size(); ++i) s += v[i] * (1.0 + (i % 7));
    return s;
}
int main() {
    std::mt19937 rng(9999);
    std::normal_distribution<double> nd;
    cublasHandle_t h;
    cublasCreate(&h);
    const double one = 1.0, zero = 0.0, half = 0.5;
    int shapes[4][3] = {{64, 48, 80}, {64, 48, 80}, {33, 47, 45}, {64, 48, 80}};
    std::vector<double> hA(128 * 128), hB(128 * 128);
    for (auto &x : hA) x = nd(rng);
    for (auto &x : hB) x = nd(rng);
    double *A, *B, *C;
    cudaMalloc(&A, sizeof(double) * 128 * 128), cudaMalloc(&B, sizeof(double) * 128 * 128), cudaMalloc(&C, sizeof(double) * 128 * 128);
    cudaMemcpy(A, hA.data(), sizeof(double) * hA.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(B, hB.data(), sizeof(double) * hB.size(), cudaMemcpyHostToDevice);
    for (int it = 0; it < 4; ++it) {
        const int m = shapes[it][0], n = shapes[it][1], k = shapes[it][2];
        if (it == 3) { // overwrite B in place: the hook cannot know (pointer identity) unless the caller disables skipping
            for (auto &x : hB) x = nd(rng);
            cudaMemcpy(B, hB.data(), sizeof(double) * hB.size(), cudaMemcpyHostToDevice);
        }
        cudaMemset(C, 0, sizeof(double) * 128 * 128);
        cublasDgemm(h, CUBLAS_OP_N, it == 2 ? CUBLAS_OP_T : CUBLAS_OP_N, m, n, k, it == 1 ? &half : &one, A, 128, B, 128, &zero, C, 128);
        std::vector<double> hC(128 * 128);
        cudaMemcpy(hC.data(), C, sizeof(double) * hC.size(), cudaMemcpyDeviceToHost);
        std::printf("dgemm %d %dx%dx%d checksum %.17g\n", it, m, n, k, checksum(hC));
    }
    // float path
    std::vector<float> fA(64 * 64), fB(64 * 64), fC(64 * 64);
    for (auto &x : fA) x = (float)nd(rng);
    for (auto &x : fB) x = (float)nd(rng);
    float *dA, *dB, *dC;
    cudaMalloc(&dA, 4 * 64 * 64), cudaMalloc(&dB, 4 * 64 * 64), cudaMalloc(&dC, 4 * 64 * 64);
    cudaMemcpy(dA, fA.data(), 4 * 64 * 64, cudaMemcpyHostToDevice), cudaMemcpy(dB, fB.data(), 4 * 64 * 64, cudaMemcpyHostToDevice);
    const float fone = 1.f, fzero = 0.f;
    cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, 64, 64, 64, &fone, dA, 64, dB, 64, &fzero, dC, 64);
    cudaMemcpy(fC.data(), dC, 4 * 64 * 64, cudaMemcpyDeviceToHost);
    double s = 0;
    for (size_t i = 0; i < fC.size(); ++i) s += fC[i] * (1.0 + (i % 7));
    std::printf("sgemm 64 checksum %.9g\n", s);
    cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, 64, 48, 80, &one, A, CUDA_R_64F, 128, B, CUDA_R_64F, 128, &zero, C, CUDA_R_64F, 128,
                 CUBLAS_COMPUTE_64F, CUBLAS_GEMM_DEFAULT);
    std::vector<double> hC(128 * 128);
    cudaMemcpy(hC.data(), C, sizeof(double) * hC.size(), cudaMemcpyDeviceToHost);
    std::printf("gemmex 64x48x80 checksum %.17g\n", checksum(hC));
    cublasDestroy(h);
    return 0;
}
