// Restated from the reference's sample programs (sample/dgemm_cuBLAS{,Lt}_int8.cu): the 4x5 * 5x3 known-answer DGEMM through
// gemmul8::gemmLt and gemmul8::gemm of OUR libgemmul8, compiled against OUR include/gemmul8.hpp.  Exit code 0 iff exact.
#include "../../include/gemmul8.hpp"
#include <cstdio>
#include <vector>
int main() {
    const int m = 4, n = 3, k = 5;
    std::vector<double> hA = {0x1.13491b78f7ff1p-1, 0x1.d5797d024f750p+0, -0x1.2121e4d9576a2p+1, 0x1.b96ec80cedfb6p-1, 0x1.466a65212f053p-2,
                              -0x1.4ec4a901fe3c4p+0, -0x1.bbff8c0e700a1p-2, 0x1.5ed8f2ba5f2dbp-2, 0x1.ca08e9321d439p+1, 0x1.627ce99fd7ed1p+1,
                              -0x1.599230c5450f8p+0, 0x1.84785f44e10f1p+1, 0x1.73682ebd0c291p-1, -0x1.0245d3a33f7d8p-4, 0x1.6df2c829f659fp-1,
                              -0x1.a3c53ea980203p-3, -0x1.fc7ec8b9281f7p-4, 0x1.7d5cd28a5e35bp+0, 0x1.68b67bfca10cfp+0, 0x1.6acd1f3bd1cafp+0};
    std::vector<double> hB = {0x1.57ce78e868ad7p-1, -0x1.351ddceb47a8bp+0, 0x1.6f39e78dc4de4p-1, 0x1.a1571993bf63bp+0, 0x1.f4a0918ad43eep-2,
                              0x1.08e1a41eff3c4p+0, 0x1.742a49c7a8c1fp-1, -0x1.36b937c0e54f0p-2, 0x1.2ceca451a1789p-2, -0x1.9316bb4db16cfp-1,
                              0x1.c6dbcad09ddd8p-1, -0x1.25a662f3a6d75p+0, -0x1.11a17e8d7e02fp+0, -0x1.9e769ce56b489p-1, -0x1.78de4dacf30d6p+1};
    std::vector<double> hX = {0x1.d51136ef01e9dp+1, 0x1.5b07528da2db2p+2, -0x1.b7d034d197c42p-4, 0x1.59b0e0e988db5p+1, 0x1.ad784e3b16dc5p-7,
                              -0x1.15b1323003b06p+0, -0x1.922e5c1c4b38bp+1, -0x1.e95843f74c224p-1, -0x1.f79e85fefa19bp+1, -0x1.0a9fa599dc6d9p+2,
                              -0x1.32cc3fa2fc921p+2, -0x1.b82c3fad3ab16p+2};
    double *A, *B, *C;
    cudaMalloc(&A, sizeof(double) * m * k), cudaMalloc(&B, sizeof(double) * k * n), cudaMalloc(&C, sizeof(double) * m * n);
    cudaMemcpy(A, hA.data(), sizeof(double) * m * k, cudaMemcpyHostToDevice);
    cudaMemcpy(B, hB.data(), sizeof(double) * k * n, cudaMemcpyHostToDevice);
    void *work;
    cudaMalloc(&work, gemmul8::workSize<false, gemmul8::Backend::INT8>(m, n, k, 15));
    const double one = 1.0, zero = 0.0;
    int bad = 0;
    for (int which = 0; which < 2; ++which) {
        cudaMemset(C, 0, sizeof(double) * m * n);
        if (which == 0) {
            cublasLtHandle_t lt;
            cublasLtCreate(&lt);
            gemmul8::gemmLt<double, gemmul8::Backend::INT8>(lt, CUBLAS_OP_N, CUBLAS_OP_N, m, n, k, &one, A, m, B, k, &zero, C, m, 15, false, work);
            cublasLtDestroy(lt);
        } else {
            cublasHandle_t h;
            cublasCreate(&h);
            gemmul8::gemm<double>(h, CUBLAS_OP_N, CUBLAS_OP_N, m, n, k, &one, A, m, B, k, &zero, C, m, 15, false, work);
            cublasDestroy(h);
        }
        std::vector<double> hC(m * n);
        cudaMemcpy(hC.data(), C, sizeof(double) * m * n, cudaMemcpyDeviceToHost);
        for (int i = 0; i < m * n; ++i) bad += hC[i] != hX[i];
    }
    std::printf("sample_kat: %s\n", bad ? "MISMATCH" : "exact");
    return bad != 0;
}
