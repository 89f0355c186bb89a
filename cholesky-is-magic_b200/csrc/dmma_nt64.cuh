// Update variant of the FP64 tensor-core NT contraction (dmma_nt.cuh):  C(128 x 64 half tiles) += alpha X Y'
// with TWO CTAs PER SM.
//
// The Cholesky trailing updates have a short contraction (K = 128 .. 512 = 4 .. 16 k-chunks), so with the
// 128 x 128 tile / one-CTA-per-SM kernel the read-modify-write epilogue of a tile (256 KB through L2) and
// the pipeline fill of the next one are exposed: the DMMA pipe sits at ~68% (K = 256) / ~80% (K = 512) of
// the formation kernel's rate (ncu launch lists in profiles/).  Here a work item is one 128 x 64 half of a
// 128 x 128 tile, computed by 4 warps (2 x 2, the same 64 x 32 warp tile and fragment layout as dmma_nt)
// with 128 threads x 254 registers and a 2-stage operand pipeline (103 KB), so two CTAs are resident on an
// SM and one's epilogue and prologue overlap the other's mainloop -- the arrangement mf_syrk_kernel uses in
// the sparse path.  Operands: X rows by 132 x 32 TMA boxes, Y rows by 68 x 32 boxes (pitches = 4 mod 16:
// conflict-free m8n8k4 fragment loads); on diagonal tiles of a symmetric product the Y fragment is read
// from the X tile.  No scale, no split-k, no batch mode: this is the alpha = -1, beta = 1 update only.
#pragma once
#include "dmma_nt.cuh"

namespace nes {

constexpr int N64_BN = 64;
constexpr int N64_PITCH_B = 68;
constexpr int N64_STAGES = 2;
constexpr int N64_WARPS = 4;
constexpr int N64_THREADS = N64_WARPS * 32;
constexpr int N64_TILE_A = NT_PITCH * NT_BK * 8;      // 33792
constexpr int N64_TILE_B = N64_PITCH_B * NT_BK * 8;   // 17408
constexpr int N64_STAGE_BYTES = N64_TILE_A + N64_TILE_B;
constexpr int N64_SMEM_BYTES = N64_STAGES * N64_STAGE_BYTES + 2 * N64_STAGES * 8 + 128;

// item -> (128 x 128 tile, half); returns false when the half lies beyond the region's columns
__device__ __forceinline__ bool n64_item(const NtArgs& p, int item, int& bi, int& bj, int& h) {
    nt_tile_coords(p, item >> 1, bi, bj);
    h = item & 1;
    return bj * NT_BN + h * N64_BN < p.N;
}

__global__ void __launch_bounds__(N64_THREADS, 2)
dmma_nt64_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
                 const NtArgs p) {
    // generic addressing of the staging buffers, as in dmma_nt_kernel (a shared-typed pointer -- LDS -- was
    // tried: 366 instead of 354 ms per factorization at m = 32768, and the lost-update defect described in
    // DESIGN.md section 5 became rarer, 1 in 5 runs instead of every run, but did not go away)
    extern __shared__ uint8_t smem_raw64[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw64) + 127) & ~static_cast<uintptr_t>(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + N64_STAGES * N64_STAGE_BYTES);
    uint64_t* empty = full + N64_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_producer = (threadIdx.x == 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < N64_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], N64_WARPS);
        }
        fence_mbar_init();
        prefetch_tmap(&mapX);
        prefetch_tmap(&mapY);
    }
    __syncthreads();

    const int kchunks = (p.K + NT_BK - 1) / NT_BK;
    const uint32_t never = (p.K < 0) ? 0xffffffffu : 0u;  // always 0 (K > 0), unknown to the compiler
    const int nitems = 2 * p.ntiles;

    // producer state (thread 0): next chunk to load
    int p_item = blockIdx.x, p_kc = 0, p_bi = 0, p_bj = 0, p_h = 0;
    uint32_t p_it = 0;
    bool p_valid = false;
    auto p_advance = [&]() {
        while (p_item < nitems && !n64_item(p, p_item, p_bi, p_bj, p_h)) p_item += gridDim.x;
        p_valid = p_item < nitems;
        p_kc = 0;
    };
    auto produce = [&]() {
        if (!p_valid) return;
        const int s = p_it % N64_STAGES;
        const uint32_t ph = (p_it / N64_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const bool diag = p.same_operand && (p_bi == p_bj);
        uint8_t* st = smem + s * N64_STAGE_BYTES;
        mbar_expect_tx(&full[s], diag ? N64_TILE_A : N64_TILE_A + N64_TILE_B);
        const int kk = p.k0 + p_kc * NT_BK;
        tma_load_2d(st, &mapX, p.rowA0 + p_bi * NT_BM, kk, &full[s]);
        if (!diag) tma_load_2d(st + N64_TILE_A, &mapY, p.rowB0 + p_bj * NT_BN + p_h * N64_BN, kk, &full[s]);
        ++p_it;
        if (++p_kc >= kchunks) {
            p_item += gridDim.x;
            p_advance();
        }
    };
    if (is_producer) {
        p_advance();
        for (int i = 0; i < N64_STAGES - 1; ++i) produce();
    }

    const int g = lane >> 2, t4 = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int a_off = t4 * NT_PITCH + wm * 64 + g;

    uint32_t it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        int bi, bj, h;
        if (!n64_item(p, item, bi, bj, h)) continue;
        const bool diag = p.same_operand && (bi == bj);
        const int pb = diag ? NT_PITCH : N64_PITCH_B;
        const int b_off = t4 * pb + (diag ? h * N64_BN : 0) + wn * 32 + g;
        const int row_base = bi * NT_BM + wm * 64 + g;
        const int col_base = bj * NT_BN + h * N64_BN + wn * 32 + 2 * t4;

        // warm L2 with the C half tile while the mainloop runs
        if (p.beta != 0.0 && g == 0 && p.batch_rows == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int col = col_base + j * 8 + cc;
                    if (col < p.N) {
                        const double* cp = p.C + (long long)col * p.ldc;
#pragma unroll
                        for (int i = 0; i < 8; i += 2)
                            if (row_base + i * 8 < p.M) prefetch_l2(cp + row_base + i * 8);
                    }
                }
        }

        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kc = 0; kc < kchunks; ++kc, ++it) {
            if (is_producer) produce();
            __syncwarp();
            const int s = it % N64_STAGES;
            const uint32_t ph = (it / N64_STAGES) & 1;
            mbar_wait(&full[s], ph);
            const double* sA = reinterpret_cast<const double*>(smem + s * N64_STAGE_BYTES);
            const double* sB = diag ? sA : sA + NT_PITCH * NT_BK;
            const double* ap = sA + a_off;
            const double* bp = sB + b_off;
            uint32_t dep = 0;
#pragma unroll
            for (int ks = 0; ks < NT_BK / 4; ++ks) {
                double af[8], bf[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) af[i] = ap[ks * 4 * NT_PITCH + i * 8];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = bp[ks * 4 * pb + j * 8];
                dep = frag_dependency(dep, af, bf);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
            // the stage is released only once every fragment load of it has returned (ptx_util.cuh)
            if (lane == 0) mbar_arrive_after(&empty[s], dep, never);
        }

        // epilogue: lane holds rows g (+8i), cols 2*t4, 2*t4+1 of each 8x8 block; a column pair is loaded
        // (16 values) before it is stored so the loads overlap
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double old[2][8];
            if (p.beta != 0.0) {
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int col = col_base + j * 8 + cc;
                    const double* cp = p.C + (long long)col * p.ldc;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = row_base + i * 8;
                        old[cc][i] = (col < p.N && row < p.M) ? __ldcg(cp + row) : 0.0;
                    }
                }
            }
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int col = col_base + j * 8 + cc;
                if (col < p.N) {
                    double* cp = p.C + (long long)col * p.ldc;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = row_base + i * 8;
                        if (row < p.M) {
                            double v = p.alpha * acc[i][j][cc];
                            if (p.beta != 0.0) v = fma(p.beta, old[cc][i], v);
                            cp[row] = v;
                        }
                    }
                }
            }
        }
    }
}

inline cudaError_t n64_configure() {
    static PerDeviceOnce once;
    int dev;
    if (!once.begin(&dev)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(dmma_nt64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         131072);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(dmma_nt64_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
    once.finish(dev, e == cudaSuccess);
    return e;
}

// mapX: 132 x 32 boxes over X, mapY: 68 x 32 boxes over Y.  max_ctas = 0: one half tile per CTA (SMs come
// free at half-tile granularity for higher-priority streams), else a persistent grid of that many CTAs.
inline cudaError_t nt64_launch(const CUtensorMap& mapX, const CUtensorMap& mapY, NtArgs a, int max_ctas,
                               cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0 || a.K <= 0) return cudaSuccess;
    cudaError_t e = n64_configure();
    if (e != cudaSuccess) return e;
    const int tm = (a.M + NT_BM - 1) / NT_BM, tn = (a.N + NT_BN - 1) / NT_BN;
    a.tiles_n = tn;
    if (!a.tile_list) a.ntiles = a.lower ? tm * (tm + 1) / 2 : tm * tn;
    if (a.ntiles <= 0) return cudaSuccess;
    a.batch_tiles = 0;
    a.batch_rows = getenv("NES_N64_NOPREFETCH") ? 1 : 0;  // debugging switch (the field is unused by this kernel)
    a.split_r = a.split_s = a.split_all = 0;
    a.scale = nullptr;
    const int nitems = 2 * a.ntiles;
    const int grid = (max_ctas > 0 && max_ctas < nitems) ? max_ctas : nitems;
    // NES_N64_ONE_CTA (debugging): ask for 128 KB so that only one CTA fits on an SM
    const int smem = getenv("NES_N64_ONE_CTA") ? 131072 : N64_SMEM_BYTES;
    dmma_nt64_kernel<<<grid, N64_THREADS, smem, stream>>>(mapX, mapY, a);
    return cudaGetLastError();
}

}  // namespace nes
