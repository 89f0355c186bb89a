// Dependent-chain latency probe (single warp) for the FP64 instructions the panel kernels lean on.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void lat(double* out, long long* cyc, double seed) {
    double x = seed + threadIdx.x * 1e-3, y = 1.0000001, z = 1e-9;
    long long t0, t1;
    // DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, y, z);
    t1 = clock64(); cyc[0] = t1 - t0;
    // DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x * y;
    t1 = clock64(); cyc[1] = t1 - t0;
    // rsqrt
    x = fabs(x) + 1.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.0;
    t1 = clock64(); cyc[2] = t1 - t0;
    // sqrt
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = sqrt(x) + 1.0;
    t1 = clock64(); cyc[3] = t1 - t0;
    // div
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = 1.0 / x + 1.5;
    t1 = clock64(); cyc[4] = t1 - t0;
    // shfl double
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    t1 = clock64(); cyc[5] = t1 - t0;
    // DMMA chain
    double c0 = x, c1 = z;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(y), "d"(z));
    t1 = clock64(); cyc[6] = t1 - t0;
    // LDS chain
    __shared__ double sm[64];
    sm[threadIdx.x] = (double)((threadIdx.x + 1) & 31);
    __syncwarp();
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = (int)sm[idx];
    t1 = clock64(); cyc[7] = t1 - t0;
    // MUFU.RSQ64H approx + 2 newton (custom fast rsqrt)
    x = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        double r = (double)rsqrtf((float)x);
        double e = fma(-x * r, r, 1.0);
        r = fma(r * e, 0.5 + 0.375 * e, r);
        x = r + 1.0;
    }
    t1 = clock64(); cyc[8] = t1 - t0;
    out[threadIdx.x] = x + c0 + c1 + idx;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
    lat<<<1, 32>>>(out, cyc, 1.0);
    lat<<<1, 32>>>(out, cyc, 1.0);
    long long h[16]; cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    const char* names[] = {"DFMA", "DMUL", "rsqrt(double)+add", "sqrt(double)+add", "1/x + add", "shfl(double)", "DMMA.8x8x4", "LDS.64->cvt chain", "fast rsqrt (f32 seed + newton)+add"};
    for (int i = 0; i < 9; ++i) printf("%-36s %.1f cycles/op\n", names[i], (double)h[i] / N);
    return 0;
}
