// W = L^-1 for one lower-triangular block (<= 128 x 128) held in shared memory, shared by the dense
// (dense_chol.cu: trtri_diag_kernel) and the multifrontal path (sparse_chol.cu: mf_trtri_kernel).  The
// inverses serve the SOLVE phases only (a block solve becomes a triangular matvec).
//
// One array holds both triangles: L strictly below the diagonal at S[r + c*P] (r > c) and W' on and
// above it, W(r, j) at S[j + r*P] (r >= j).  Column j of W is the forward substitution
//     W(j,j) = 1/L(j,j),   W(r,j) = -(sum_{c=j}^{r-1} L(r,c) W(c,j)) / L(r,r)
// and the columns are independent.  FOUR adjacent lanes share a column: lane h takes the terms with
// c == j + h (mod 4) and the partial sums meet in two shuffles, so the dependent chain per row is a
// quarter of the dot product (one thread per column took ~150 us per block: the longest column is a chain
// of 8 000 dependent shared-memory FMAs).  TRTRI_THREADS = 4 * 128 threads per CTA; no CTA-wide barrier
// inside (the lanes of a column sit in one warp).
#pragma once
#include <cuda_runtime.h>

namespace nes {

constexpr int TRTRI_LANES = 4;
constexpr int TRTRI_THREADS = TRTRI_LANES * 128;

template <int P>
__device__ __forceinline__ void trtri_columns_smem(double* S, const double* dv, int nc) {
    const int j = threadIdx.x / TRTRI_LANES, h = threadIdx.x % TRTRI_LANES;
    // a warp holds 8 columns; its lanes run the same number of rows (up to the end of the block), lanes of
    // finished / absent columns idle inside the loop so that the shuffles stay convergent
    const int jw = (threadIdx.x & ~31) / TRTRI_LANES;  // first column of this warp
    if (jw >= nc) return;
    double* w = S + j;                                 // w[c * P] = W(c, j)
    if (j < nc && h == 0) w[j * P] = dv[j];
    __syncwarp();
    for (int r = jw + 1; r < nc; ++r) {
        double s0 = 0.0;
        if (j < nc && r > j) {
            for (int c = j + h; c < r; c += TRTRI_LANES) s0 = fma(S[r + c * P], w[c * P], s0);
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        if (j < nc && r > j && h == 0) w[r * P] = -s0 * dv[r];
        __syncwarp();
    }
}

}  // namespace nes
