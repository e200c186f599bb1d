// corr_kernels.cu -- pairwise-complete second moments of sampled trajectory columns (sm_100a).
//
// est_scaling_params.calc_theta_explicit_ss_bounds (EST.py:161-189; the _ppd variant :211-243) turns the r1 (or dVNs)
// trajectories into a DataFrame with zeros replaced by NaN and calls DataFrame.corr(): for every pair of sampled steps
// (i, j) the Pearson correlation over the frames in which BOTH are non-zero.  With X[f][i] the sampled integer values,
//     N_ij = #{f : X_fi != 0, X_fj != 0}   Sx_ij = sum_f X_fi [X_fj != 0]   Sxx_ij = sum_f X_fi^2 [X_fj != 0]   Sxy_ij = sum_f X_fi X_fj
//     corr_ij = (N Sxy - Sx_ij Sx_ji) / sqrt((N Sxx_ij - Sx_ij^2) (N Sxx_ji - Sx_ji^2))
// All four are exact int64 sums, additive over batches and ranks (one all-reduce), so 10^5..10^7 trajectories reduce to
// 4 K x K matrices on the device instead of a [frames][K] float table on the host.  The contraction is a masked X^T X;
// integer exactness (values up to 2.5 M at M = 10^5, sums beyond 2^53) rules out the bf16/tf32 tensor-core path, and at
// K <= 1000 sampled steps the IMAD.WIDE loop below finishes 10^5 frames in tens of milliseconds.
#include "common.cuh"

namespace scldpc {

#define CORR_TILE 32

// grid (ceil(K/32), ceil(K/32)); block (32, 32): thread (tx, ty) owns pair (i = bx*32+ty, j = by*32+tx)
__global__ void __launch_bounds__(CORR_TILE *CORR_TILE) corr_moments_kernel(const int32_t *r1, int n_frames, int row_len, int start,
                                                                            int step, int K, long long *acc)
{
    __shared__ int xi[CORR_TILE][CORR_TILE + 1], xj[CORR_TILE][CORR_TILE + 1];     // [frame in tile][column in tile]
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * CORR_TILE + ty, j = blockIdx.y * CORR_TILE + tx;
    long long n = 0, sx = 0, sxx = 0, sxy = 0;
    for (int f0 = 0; f0 < n_frames; f0 += CORR_TILE) {
        // thread (tx, ty) loads frame f0+ty, columns bx*32+tx and by*32+tx
        const int f = f0 + ty;
        const int ci = blockIdx.x * CORR_TILE + tx, cj = blockIdx.y * CORR_TILE + tx;
        xi[ty][tx] = (f < n_frames && ci < K) ? r1[(size_t)f * row_len + start + (size_t)ci * step] : 0;
        xj[ty][tx] = (f < n_frames && cj < K) ? r1[(size_t)f * row_len + start + (size_t)cj * step] : 0;
        __syncthreads();
#pragma unroll 8
        for (int q = 0; q < CORR_TILE; q++) {
            const long long a = xi[q][ty], b = xj[q][tx];
            const long long mb = (b != 0);
            n += (a != 0) & mb;
            sx += a * mb;
            sxx += a * a * mb;
            sxy += a * b;
        }
        __syncthreads();
    }
    if (i < K && j < K) {
        const size_t KK = (size_t)K * K, o = (size_t)i * K + j;
        acc[o] += n; acc[KK + o] += sx; acc[2 * KK + o] += sxx; acc[3 * KK + o] += sxy;
    }
}

void corr_moments_launch(const int32_t *r1, int n_frames, int row_len, int start, int step, int K, long long *acc, cudaStream_t st)
{
    const int t = (K + CORR_TILE - 1) / CORR_TILE;
    g_prof.launches += 1;
    corr_moments_kernel<<<dim3(t, t), dim3(CORR_TILE, CORR_TILE), 0, st>>>(r1, n_frames, row_len, start, step, K, acc);
}

}  // namespace scldpc
