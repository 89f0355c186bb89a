// Probe: cp.async.bulk.tensor with the tensor map in (a) kernel param space, (b) global memory, for a
// large and a tiny tensor (box larger than the tensor).  Usage: tma_gmem_probe <case 0..5>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../cholesky-is-magic_b200/csrc/dmma_nt.cuh"
using namespace nes;

__global__ void probe_param(const __grid_constant__ CUtensorMap map, double* out, int fence, int c0) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + NT_TILE_BYTES);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, NT_TILE_BYTES);
        tma_load_2d(sm, &map, c0, 0, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    const double* s = reinterpret_cast<const double*>(sm);
    if (threadIdx.x < 8) out[threadIdx.x] = s[threadIdx.x] + 10.0 * s[NT_PITCH + threadIdx.x];
}
__global__ void probe_global(const CUtensorMap* maps, int which, double* out, int fence, int c0) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + NT_TILE_BYTES);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, NT_TILE_BYTES);
        if (fence) fence_tensormap_acquire(maps + which);
        tma_load_2d(sm, maps + which, c0, 0, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    const double* s = reinterpret_cast<const double*>(sm);
    if (threadIdx.x < 8) out[threadIdx.x] = s[threadIdx.x] + 10.0 * s[NT_PITCH + threadIdx.x];
}
int main(int argc, char** argv) {
    const int cs = argc > 1 ? atoi(argv[1]) : 0;
    const bool tiny = cs & 1;
    const int where = cs >> 1;  // 0 param, 1 global, 2 global + fence
    int rows = tiny ? 5 : 200, cols = tiny ? 2 : 64, ld = tiny ? 16 : 208;
    const int c0 = argc > 5 ? atoi(argv[5]) : 0;
    if (argc > 4) { rows = atoi(argv[2]); cols = atoi(argv[3]); ld = atoi(argv[4]); }
    std::vector<double> h((size_t)ld * cols);
    for (int j = 0; j < cols; ++j) for (int i = 0; i < ld; ++i) h[i + (size_t)j * ld] = i + 0.01 * j;
    double *d, *out; cudaMalloc(&d, h.size() * 8 + 4096); cudaMalloc(&out, 64);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    CUtensorMap map;
    printf("c0 %d ", c0); printf("case %d: %s tensor %dx%d ld %d, map in %s: encode rc %d\n", cs, tiny ? "tiny" : "large", rows, cols, ld,
           where == 0 ? "param" : where == 1 ? "global" : "global+fence", make_operand_map(&map, d, rows, cols, ld));
    CUtensorMap* dm; cudaMalloc(&dm, 4 * sizeof(CUtensorMap));
    cudaMemcpy(dm + 1, &map, sizeof(map), cudaMemcpyHostToDevice);
    const int smem = NT_TILE_BYTES + 64;
    cudaFuncSetAttribute(probe_param, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe_global, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (where == 0) probe_param<<<1, 32, smem>>>(map, out, 0, c0);
    else probe_global<<<1, 32, smem>>>(dm, 1, out, where == 2, c0);
    cudaError_t e = cudaDeviceSynchronize();
    double r[8] = {0};
    if (e == cudaSuccess) cudaMemcpy(r, out, 64, cudaMemcpyDeviceToHost);
    printf("  -> %s; out = %.2f %.2f %.2f %.2f %.2f %.2f\n", cudaGetErrorString(e), r[0], r[1], r[2], r[3], r[4], r[5]);
    return 0;
}
