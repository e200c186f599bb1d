// traj_kernels.cu -- per-iteration moments of the BP trajectories, accumulated on the device (sm_100a).
//
// The notebook estimates the BP ("parallel peeling") scaling parameters from the per-iteration rows bp_traj prints
// (iter, deg_1_iter, dVNs, first erased position; BP_TRAJ.c:988,1051): the mean dVNs over the frames still decoding
// (NB cell 41: sum / count of non-zeros), the variance around the mean-evolution curve (NB cell 42:
// mean((dVNs/N - theory/N)^2) * N over ALL frames, finished ones padded with zeros) and, for the sliding-window law, the
// position of the first erased VN.  All of them are functions of, per iteration t,
//     frames that executed t, frames with dVNs != 0, sum dVNs, sum dVNs^2, sum deg1, sum deg1^2, sum first_pos, sum dVNs*deg1
// which are exact integers, additive over batches, ranks and runs -- the BP analogue of the peeling path's
// (ssquares, counts) pickle (PD.py:1286-1294).  With them 10^5..10^7 frames never leave the device as text rows.
#include "common.cuh"

namespace scldpc {

#define TRAJ_MOMENTS 8

// rows [G][max_rows][lanes][3] = (deg1, dVNs, first_pos), zero beyond a frame's last iteration; iters [G][lanes].
// One block per iteration t; acc [max_rows][TRAJ_MOMENTS] (+=).
__global__ void __launch_bounds__(256) traj_moments_kernel(const int32_t *rows, const int32_t *iters, int G, int max_rows, int lanes,
                                                           int n_frames, long long *acc)
{
    const int t = blockIdx.x;
    long long m[TRAJ_MOMENTS] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < G * lanes; i += blockDim.x) {
        const int g = i / lanes, l = i % lanes;
        if (l >= n_frames) continue;
        if (iters[g * lanes + l] <= t) continue;                 // the frame had stopped: its row is the zero padding
        const int32_t *r = rows + (((size_t)g * max_rows + t) * lanes + l) * 3;
        const long long d1 = r[0], dv = r[1], fp = r[2];
        m[0] += 1; m[1] += (dv != 0); m[2] += dv; m[3] += dv * dv; m[4] += d1; m[5] += d1 * d1; m[6] += fp; m[7] += dv * d1;
    }
    __shared__ long long s[TRAJ_MOMENTS][8];
#pragma unroll
    for (int q = 0; q < TRAJ_MOMENTS; q++) {
        long long v = m[q];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s[q][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < TRAJ_MOMENTS) {
        long long v = 0;
        for (int w = 0; w < 8; w++) v += s[threadIdx.x][w];
        acc[(size_t)t * TRAJ_MOMENTS + threadIdx.x] += v;        // one block per t: no race
    }
}

void traj_moments_launch(const int32_t *rows, const int32_t *iters, int G, int max_rows, int lanes, int n_frames, long long *acc,
                         cudaStream_t st)
{
    g_prof.launches += 1;
    traj_moments_kernel<<<max_rows, 256, 0, st>>>(rows, iters, G, max_rows, lanes, n_frames, acc);
}

}  // namespace scldpc
