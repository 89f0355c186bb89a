// K7/K8: fused elementwise maps and single-pass reductions of the interior-point drivers.
// Every kernel is one pass over n- or m-vectors (HBM/latency-bound, negligible flops); reductions
// are block partials + one fixed-order finishing kernel, so results are bitwise reproducible.
#pragma once
#include <cuda_runtime.h>

#include <cmath>

namespace nes {

enum RedOp { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };
constexpr int RED_THREADS = 256;
constexpr int RED_MAXN = 8;  // reductions fused per kernel

__device__ __forceinline__ double red_combine(double a, double b, int op) {
    return op == RED_SUM ? a + b : (op == RED_MAX ? fmax(a, b) : fmin(a, b));
}

__device__ __forceinline__ double red_identity(int op) {
    return op == RED_SUM ? 0.0 : (op == RED_MAX ? -INFINITY : INFINITY);
}

// Block-wide reduction of `v`; the result is valid in thread 0.  `buf` is 32 doubles of smem.
__device__ __forceinline__ double block_reduce(double v, int op, double* buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = red_combine(v, __shfl_xor_sync(0xffffffffu, v, off), op);
    __syncthreads();
    if (lane == 0) buf[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < (blockDim.x >> 5)) ? buf[lane] : red_identity(op);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            v = red_combine(v, __shfl_xor_sync(0xffffffffu, v, off), op);
    }
    return v;
}

// partial[k * nblocks + b], k < nred; ops packed 2 bits each.
static __global__ void reduce_finish_kernel(const double* __restrict__ partial, int nblocks, int nred,
                                     unsigned ops, double* __restrict__ out) {
    const int k = threadIdx.x;
    if (k >= nred) return;
    const int op = (ops >> (2 * k)) & 3;
    double v = red_identity(op);
    for (int b = 0; b < nblocks; ++b) v = red_combine(v, partial[k * nblocks + b], op);
    out[k] = v;
}

__host__ __device__ constexpr unsigned pack_ops(int o0 = 0, int o1 = 0, int o2 = 0, int o3 = 0, int o4 = 0,
                                                int o5 = 0, int o6 = 0, int o7 = 0) {
    return (unsigned)o0 | ((unsigned)o1 << 2) | ((unsigned)o2 << 4) | ((unsigned)o3 << 6) |
           ((unsigned)o4 << 8) | ((unsigned)o5 << 10) | ((unsigned)o6 << 12) | ((unsigned)o7 << 14);
}

}  // namespace nes
