// Vendor-library comparator for the FP64 roofline (tools only; libnes.so never links cuBLAS / cuSOLVER):
//   cublasDgemm 8192^3, cublasDsyrk (the formation's shape: n = m, k = 2m), cusolverDnDpotrf,
// next to the repo's own numbers for the same shapes (bench.py).  "Beat the vendor kernel on the same box."
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/vendor_fp64 tools/vendor_fp64.cu -lcublas -lcusolver
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void fill_kernel(double* p, size_t n, unsigned long long seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long x = (i + seed) * 0x9E3779B97F4A7C15ull;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        p[i] = (double)(x >> 11) * (1.0 / 9007199254740992.0);
    }
}
__global__ void add_diag_kernel(double* p, int m, size_t ld, double v) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) p[i + (size_t)i * ld] += v;
}

template <class F> static double time_ms(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char** argv) {
    const int m = argc > 1 ? atoi(argv[1]) : 8192;
    const int n = argc > 2 ? atoi(argv[2]) : 2 * m;
    cublasHandle_t h; cublasCreate(&h);
    cusolverDnHandle_t s; cusolverDnCreate(&s);
    double *A, *C, *B;
    CK(cudaMalloc(&A, (size_t)m * n * 8)); CK(cudaMalloc(&C, (size_t)m * m * 8)); CK(cudaMalloc(&B, (size_t)m * m * 8));
    fill_kernel<<<1024, 256>>>(A, (size_t)m * n, 1);
    fill_kernel<<<1024, 256>>>(B, (size_t)m * m, 2);
    const double one = 1.0, zero = 0.0;
    // DGEMM m x m x m (B * B')
    double ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, m, m, m, &one, B, m, B, m, &zero, C, m); }, 3);
    printf("cublasDgemm  %d^3: %.2f ms  %.2f TFLOP/s\n", m, ms, 2.0 * m * m * m / ms / 1e9);
    // DSYRK: C(lower) = A A', A m x n  (the formation's shape; algorithmic flops m^2 n)
    ms = time_ms([&] { cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, m, n, &one, A, m, &zero, C, m); }, 3);
    printf("cublasDsyrk  m=%d k=%d: %.2f ms  %.2f TFLOP/s (on m^2 k)\n", m, n, ms, (double)m * m * n / ms / 1e9);
    // DPOTRF of C + m I (fresh copy per repetition)
    add_diag_kernel<<<64, 256>>>(C, m, m, (double)n);
    int lwork = 0; cusolverDnDpotrf_bufferSize(s, CUBLAS_FILL_MODE_LOWER, m, C, m, &lwork);
    double* work; int* info; CK(cudaMalloc(&work, (size_t)lwork * 8)); CK(cudaMalloc(&info, 4));
    double best = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemcpy(B, C, (size_t)m * m * 8, cudaMemcpyDeviceToDevice));
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        cusolverDnDpotrf(s, CUBLAS_FILL_MODE_LOWER, m, B, m, work, lwork, info);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float t = 0; cudaEventElapsedTime(&t, a, b);
        if (rep > 0 && t < best) best = t;
    }
    int hinfo = -1; CK(cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost));
    printf("cusolverDnDpotrf m=%d: %.2f ms  %.2f TFLOP/s (on m^3/3)  info %d\n", m, best, (double)m * m * m / 3.0 / best / 1e9, hinfo);
    return 0;
}
