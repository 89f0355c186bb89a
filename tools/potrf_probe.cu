// phase timing of the in-smem 128x128 Cholesky (potrf_block.cuh) with clock64
#define POTRF_PROFILE
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../cholesky-is-magic_b200/csrc/potrf_block.cuh"
using namespace nes;
__global__ void __launch_bounds__(256) k(double* M, int jb, double* dinv_out, int* info) {
    extern __shared__ __align__(128) double S[];
    double* dinv = S + CH_NB * CH_P;
    for (int idx = threadIdx.x; idx < jb * jb; idx += 256) S[idx] = M[idx];
    __syncthreads();
    long long t0 = clock64();
    potrf_block_smem(S, dinv, jb, 0.0, info, 0);
    long long t1 = clock64();
    for (int idx = threadIdx.x; idx < jb * jb; idx += 256) M[idx] = S[idx];
    if (threadIdx.x == 0) { dinv_out[0] = (double)(t1 - t0); }
}
int main() {
    const int n = 128;
    std::vector<double> h(n * n), B(n * n);
    for (int i = 0; i < n * n; ++i) B[i] = (double)((i * 2654435761u) % 1000) / 1000.0;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = (i == j) ? n : 0; for (int k = 0; k < n; ++k) s += B[i + k * n] * B[j + k * n]; h[i + j * n] = s; }
    double *d, *dv; int* info;
    cudaMalloc(&d, n * n * 8); cudaMalloc(&dv, 1024); cudaMalloc(&info, 16); cudaMemset(info, 0, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (n * n + n) * 8);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(d, h.data(), n * n * 8, cudaMemcpyHostToDevice);
        k<<<1, 256, (n * n + n) * 8>>>(d, n, dv, info);
        cudaDeviceSynchronize();
        double tot; long long prof[8];
        cudaMemcpy(&tot, dv, 8, cudaMemcpyDeviceToHost);
        cudaMemcpyFromSymbol(prof, potrf_prof, 64);
        printf("total %.0f cycles, pivot+rows %lld, update %lld | first sub-panel: loads %lld factor %lld solve+store %lld, phase %lld t96 %lld t255 %lld (%s)\n", tot, prof[0], prof[1], prof[2], prof[3], prof[4], prof[5], prof[6], prof[7], cudaGetErrorString(cudaGetLastError()));
    }
    std::vector<double> L(n * n); cudaMemcpy(L.data(), d, n * n * 8, cudaMemcpyDeviceToHost);
    double err = 0, nrm = 0;
    for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k2 = 0; k2 <= j; ++k2) s += L[i + k2 * n] * L[j + k2 * n]; err += (s - h[i + j * n]) * (s - h[i + j * n]); nrm += h[i + j * n] * h[i + j * n]; }
    printf("residual %.2e\n", sqrt(err / nrm));
    return 0;
}
