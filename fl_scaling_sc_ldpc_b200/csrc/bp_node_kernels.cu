// bp_node_kernels.cu -- frame streams in the node-state formulation of flooding BP on the BEC (sm_100a).
//
// From the all-erased start the messages of flooding BP on the BEC are monotone, and the a-posteriori erased set after
// iteration t is a function of the erased set after iteration t-1 alone ("parallel peeling"):
//
//     x_v(t) = x_v(t-1)  AND NOT  (some CN of v has v as its only erased neighbour in x(t-1))
//
// (an extrinsic message that the message-passing decoder would still hold back can only concern a VN that is already
// known).  The erased set -- hence the residual erasures, the per-frame iteration count under the
// reference's stall rule (NumErasures == NumErasuresPrec, BP_FULL.c:1046-1066), the error counts and the expurgation
// inputs -- is bit-identical to decodeBP's at every iteration; tests/test_stream_gpu.py holds it against the message
// kernels (themselves checked against the oracle and the compiled reference).  What the formulation does not carry are
// the messages, so the trajectory mode (deg_1_iter, BP_TRAJ.c:935-979) stays on the message kernels (bp_kernels.cu /
// bp_wave_kernels.cu), which remain the implementation of record; the window decoder's node-state form is in bp_kernels.cu.
//
// State per graph: x and xb [n][chunks] (1 bit per VN and frame; equal between iterations).
//   CN sweep  : gathers the dc x rows of a CN (E rows through L2: the gather stays inside a band of dv positions, 5 MB at
//               M = 10000 and 1024 frames) and clears, in xb, the neighbour it resolves (sparse 32-bit atomics)
//   state pass: sequential; copies the touched rows of xb to x, ORs "an erased VN is left", arms freed lanes
// HBM sees the first touch of every x row, the index stream and the sequential state pass: about 2n/8 bytes per
// frame-iteration (158 KB measured against 148 KB) instead of (4E + n)/8 = 1062 KB.
//
// Useful work is still accounted as the reference's: 2E edge updates per frame-iteration.
#include <cstdlib>

#include "common.cuh"

namespace scldpc {

constexpr int NS_X_ROWS = 4;

// ------------------------------------------------------------------------------------------------------------
// check-node sweep: a CN with exactly one erased neighbour (in x, the state after the previous iteration) resolves it --
// the bit is cleared in xb, the copy that becomes the state after this iteration, so every CN of the sweep still reads the
// old state (flooding).  Resolutions are sparse (a VN is resolved once per frame), so the scatter costs little.
// The sweep is bound by L2 latency and instruction issue, not by HBM.  Measured and dropped: a cp.async pipeline for the
// index rows (no gain, +25 % instructions), a register prefetch of the next index row (-8 %), five blocks per SM at 48
// registers (-2 %).  The index of the neighbour to clear is carried in bit planes next to the saturating count, which
// keeps the divergent scatter at ~30 instructions per resolution.
// ------------------------------------------------------------------------------------------------------------
// CAPPED (streams with an iteration cap): a frame that hit the cap is not at a fixed point, so its bits must not be cleared
// while it waits for the harvest -- the resolutions are masked with the frames still iterating.
template <int DV, int DC, bool CAPPED>
__global__ void __launch_bounds__(256, 4) ns_cn_kernel(BpParams p)
{
    static_assert(DC <= 16, "neighbour index is encoded in four bit planes");
    pdl_wait_then_release();
    const int g = blockIdx.y;
    if (ld_cg(p.alive + g) == 0) return;
    __shared__ u64 s_new[SCLDPC_MAX_WORDS];
    if (threadIdx.x < SCLDPC_MAX_WORDS) s_new[threadIdx.x] = 0;
    __syncthreads();
    const int ch = p.chunks;
    const int k = threadIdx.x & (ch - 1);
    const u128 act = reinterpret_cast<const u128 *>(p.active)[g * ch + k];
    const bool lane_work = nz(act);                             // a thread keeps its chunk
    const u128 *__restrict__ xk = p.x + (size_t)g * p.n * ch + k;
    unsigned *__restrict__ xbk = reinterpret_cast<unsigned *>(p.xb + (size_t)g * p.n * ch + k);
    unsigned char *__restrict__ dirtyk = p.dirty + (size_t)g * p.n * ch + k;
    const int32_t *__restrict__ cn_edge = p.cn_edge + (size_t)g * p.nk * DC;
    const int items = p.c1 << p.chunk_shift;                    // CNs >= c1 (tail of a truncated code) are never swept
    const int stride = gridDim.x * blockDim.x;
    const int E = p.E;
    u128 acc_new = zero128();
    if (lane_work)
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < items; idx += stride) {
            const int32_t *row = cn_edge + (size_t)(idx >> p.chunk_shift) * DC;
            int e[DC];
            load_row<DC>(row, e);
            u128 in[DC];
#pragma unroll
            for (int j = 0; j < DC; j++) in[j] = (e[j] != E) ? ld_stream(xk + (unsigned)((e[j] / DV) << p.chunk_shift)) : zero128();
            // saturating count of erased neighbours (one / two planes) and, in bit planes b0..b3, the index j of an erased
            // neighbour -- exact where it is needed, i.e. in the frames with exactly one
            u128 one = zero128(), tw = zero128(), b0 = zero128(), b1 = zero128(), b2 = zero128(), b3 = zero128();
#pragma unroll
            for (int j = 0; j < DC; j++) {
                tw |= one & in[j];
                one |= in[j];
                if (j & 1) b0 |= in[j];
                if (j & 2) b1 |= in[j];
                if (j & 4) b2 |= in[j];
                if (j & 8) b3 |= in[j];
            }
            u128 res = one & ~tw;                               // frames in which exactly one neighbour of c is erased
            if (CAPPED) res &= act;
            if (nz(res)) {
                acc_new |= res;
                const unsigned rw[4] = {(unsigned)res.x, (unsigned)(res.x >> 32), (unsigned)res.y, (unsigned)(res.y >> 32)};
                const unsigned w0[4] = {(unsigned)b0.x, (unsigned)(b0.x >> 32), (unsigned)b0.y, (unsigned)(b0.y >> 32)};
                const unsigned w1[4] = {(unsigned)b1.x, (unsigned)(b1.x >> 32), (unsigned)b1.y, (unsigned)(b1.y >> 32)};
                const unsigned w2[4] = {(unsigned)b2.x, (unsigned)(b2.x >> 32), (unsigned)b2.y, (unsigned)(b2.y >> 32)};
                const unsigned w3[4] = {(unsigned)b3.x, (unsigned)(b3.x >> 32), (unsigned)b3.y, (unsigned)(b3.y >> 32)};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    unsigned m = rw[q];
                    while (m) {
                        const int b = __ffs((int)m) - 1;
                        m &= m - 1;
                        int j = ((w0[q] >> b) & 1u) | (((w1[q] >> b) & 1u) << 1) | (((w2[q] >> b) & 1u) << 2);
                        if (DC > 8) j |= ((w3[q] >> b) & 1u) << 3;
                        const unsigned o = (unsigned)((__ldg(row + j) / DV) << p.chunk_shift);    // L1 hit
                        atomicAnd(xbk + 4 * (size_t)o + q, ~(1u << b));
                        dirtyk[o] = 1;
                    }
                }
            }
        }
    acc_new = warp_or_same_chunk(acc_new, ch);
    if ((threadIdx.x & 31) < ch) {
        if (acc_new.x) atomicOr(&s_new[2 * k], acc_new.x);
        if (acc_new.y) atomicOr(&s_new[2 * k + 1], acc_new.y);
    }
    __syncthreads();
    if (threadIdx.x < p.W) {
        const int w = threadIdx.x;
        if (s_new[w] & ~ld_cg(p.any_new + g * p.W + w)) atomicOr(p.any_new + g * p.W + w, s_new[w]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// state pass + end of the iteration (last block): brings x up to xb on the rows the CN sweep touched, collects "an erased
// VN is left", arms freed lanes with their new frames' channel draws; same control flow as bp_vn_stream_kernel
// ------------------------------------------------------------------------------------------------------------
template <bool ARM, bool CAPPED>
__global__ void __launch_bounds__(256, 4) ns_x_kernel(BpParams p)
{
    pdl_wait_then_release();
    // the CN sweep walks graphs and rows upwards, this pass downwards (vn_reverse): the CN sweep then starts on the rows
    // this pass touched last -- its first gathers are L2 hits
    const int g = p.vn_reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
    if (ld_cg(p.alive + g) == 0) return;
    __shared__ u64 s_er[SCLDPC_MAX_WORDS], s_first[SCLDPC_MAX_WORDS];
    __shared__ int s_last;
    if (threadIdx.x < SCLDPC_MAX_WORDS) { s_er[threadIdx.x] = 0; s_first[threadIdx.x] = 0; }
    __syncthreads();
    const int ch = p.chunks;
    const int k = threadIdx.x & (ch - 1);
    const u128 act = reinterpret_cast<const u128 *>(p.active)[g * ch + k];
    const u128 arm = ARM ? reinterpret_cast<const u128 *>(p.arm_mask)[g * ch + k] : zero128();
    u128 acc_er = zero128(), acc_first = zero128();
    u128 *__restrict__ x = p.x + (size_t)g * p.n * ch;
    u128 *__restrict__ xb = p.xb + (size_t)g * p.n * ch;
    unsigned char *__restrict__ dirty = p.dirty + (size_t)g * p.n * ch;
    const int items = p.n << p.chunk_shift;
    const int stride = gridDim.x * blockDim.x;
    const u64 thr = (ARM && nz(arm)) ? p.thr[g] : 0ull;
    const uint64_t gid = p.first_graph + (uint64_t)g;
    constexpr int U = NS_X_ROWS;                                // rows in flight per thread: the pass is a plain stream
    if (nz(act | arm))
        for (int base = blockIdx.x * blockDim.x * U + threadIdx.x; base < items; base += stride * U) {
            u128 xs[U];
            unsigned char ds[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int lin = base + u * (int)blockDim.x;
                const int idx = p.vn_reverse ? ((p.n - 1 - (lin >> p.chunk_shift)) << p.chunk_shift) + k : lin;
                xs[u] = zero128(); ds[u] = 0;
                if (lin < items) { xs[u] = ld_cg128(xb + idx); ds[u] = dirty[idx]; }   // the CN sweep wrote xb with atomics (L2)
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int lin = base + u * (int)blockDim.x;
                if (lin >= items) break;
                const int idx = p.vn_reverse ? ((p.n - 1 - (lin >> p.chunk_shift)) << p.chunk_shift) + k : lin;
                u128 xn = xs[u];
                const bool d = ds[u] != 0;
                bool wr_x = d, wr_b = false;
                if (ARM && nz(arm)) {
                    // new frames: the erased set starts as the channel's (Lji = channel value on every edge, BP_FULL.c:913-917)
                    const int v = idx >> p.chunk_shift;
                    const bool forced = p.known && (v % p.vns_pos) < p.known[v / p.vns_pos];
                    u128 cw = zero128();
                    // freed lanes get consecutive frame ids in ascending lane order, so one Philox call (4 frames) is
                    // usually shared by up to four armed lanes
                    uint32_t blk = 0xffffffffu, r4[4] = {0, 0, 0, 0};
                    for (int half = 0; half < 2; half++) {
                        u64 m = half ? arm.y : arm.x, w = 0;
                        while (m && !forced) {
                            const int b = __ffsll((long long)m) - 1;
                            m &= m - 1;
                            const uint32_t fr = (uint32_t)p.lane_frame[g * p.lanes + k * 128 + half * 64 + b];
                            if ((fr >> 2) != blk) {
                                blk = fr >> 2;
                                philox4x32_10((uint32_t)v, blk, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)p.seed ^ 0x6368616Eu,
                                              (uint32_t)(p.seed >> 32), r4);
                            }
                            if ((u64)r4[fr & 3] < thr) w |= 1ull << b;
                        }
                        if (half) cw.y = w; else cw.x = w;
                    }
                    const u128 xa = sel(arm, cw, xn);
                    if (neq(xa, xn)) { xn = xa; wr_x = true; wr_b = true; }
                    acc_first |= ~cw & arm;                     // NumErasuresPrec = n before the first iteration: it makes
                }                                               // "progress" iff the channel left some VN known
                if (wr_x) x[idx] = xn;
                if (wr_b) xb[idx] = xn;
                if (d) dirty[idx] = 0;
                acc_er |= xn & act;
            }
        }
    acc_er = warp_or_same_chunk(acc_er, ch);
    if (ARM) acc_first = warp_or_same_chunk(acc_first, ch);
    if ((threadIdx.x & 31) < ch) {
        if (acc_er.x) atomicOr(&s_er[2 * k], acc_er.x);
        if (acc_er.y) atomicOr(&s_er[2 * k + 1], acc_er.y);
        if (ARM) {
            if (acc_first.x) atomicOr(&s_first[2 * k], acc_first.x);
            if (acc_first.y) atomicOr(&s_first[2 * k + 1], acc_first.y);
        }
    }
    __syncthreads();
    if (threadIdx.x < p.W) {
        const int w = threadIdx.x;
        if (s_er[w] & ~ld_cg(p.any_er + g * p.W + w)) atomicOr(p.any_er + g * p.W + w, s_er[w]);
        if (ARM && (s_first[w] & ~ld_cg(p.first_new + g * p.W + w))) atomicOr(p.first_new + g * p.W + w, s_first[w]);
        __threadfence();                                        // only these words are read by the graph's last block
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(p.ticket + g, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- end of the iteration for graph g ----
    __shared__ u64 s_act[SCLDPC_MAX_WORDS], s_cap[SCLDPC_MAX_WORDS];
    const int W = p.W;
    for (int w = threadIdx.x; w < W; w += blockDim.x) { s_act[w] = p.active[g * W + w]; s_cap[w] = 0; }
    __syncthreads();
    for (int l = threadIdx.x; l < p.lanes; l += blockDim.x) {
        const int w = l >> 6, b = l & 63;
        if ((s_act[w] >> b) & 1ull) {
            const int li = p.lane_iter[g * p.lanes + l] + 1;
            p.lane_iter[g * p.lanes + l] = li;
            if (CAPPED && li >= p.stream_cap) atomicOr(reinterpret_cast<unsigned long long *>(&s_cap[w]), 1ull << b);   // while (iter < MaxNumIt)
        }
    }
    __syncthreads();
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        const u64 er = ld_cg(p.any_er + g * W + w);
        p.any_er[g * W + w] = 0;
        const u64 a = s_act[w];
        u64 nw = ld_cg(p.any_new + g * W + w);
        if (!ARM) {                                             // lanes armed by the previous pass ran their first iteration
            nw |= ld_cg(p.first_new + g * W + w);
            p.first_new[g * W + w] = 0;
        }
        const u64 stop = a & (~er | ~nw | s_cap[w]);            // NumErasures == 0  ||  == NumErasuresPrec  ||  iteration cap
        u64 left = a & ~stop;
        if (ARM) { left |= p.arm_mask[g * W + w]; p.arm_mask[g * W + w] = 0; }   // armed lanes start iterating with the next sweep
        p.active[g * W + w] = left;
        p.done_mask[g * W + w] |= stop;
        p.fail_mask[g * W + w] |= stop & er;
        p.any_new[g * W + w] = 0;
    }
    if (threadIdx.x == 0) p.ticket[g] = 0;
}

// ------------------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------------------
template <typename K>
static int resident_blocks_ns(K kernel, int block)
{
    int occ = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, 0) != cudaSuccess || occ < 1) occ = 2;
    return occ * (sms > 0 ? sms : 148);
}

template <int DV, int DC>
static void launch_node_iteration(const BpParams &p, bool arm, cudaStream_t st)
{
    const int block = 256;
    static int res_cn = 0, res_x = 0;
    if (!res_cn) {
        res_cn = resident_blocks_ns(ns_cn_kernel<DV, DC, false>, block);
        res_x = resident_blocks_ns(ns_x_kernel<true, false>, block);
    }
    // every graph gets enough blocks to fill the machine on its own: blocks of finished graphs return at once
    auto grid = [&](int resident, long long items_per_graph) {
        long long need = (items_per_graph + block - 1) / block;
        long long gx = need < resident ? need : resident;
        return dim3((unsigned)(gx < 1 ? 1 : gx), (unsigned)p.G, 1);
    };
    dim3 gc = grid(res_cn, (long long)p.c1 << p.chunk_shift);
    dim3 gx = grid(res_x, (((long long)p.n << p.chunk_shift) + NS_X_ROWS - 1) / NS_X_ROWS);
    const bool sample = g_prof.sample_every > 0 && g_prof.n_samples < g_prof.max_samples && (p.iter % g_prof.sample_every) == 0;
    cudaEvent_t *ev = sample ? g_prof.ev + 3 * g_prof.n_samples : nullptr;
    if (sample) cudaEventRecord(ev[0], st);
    g_prof.launches += 2;
    static const bool pdl = getenv("SCLDPC_NO_PDL") == nullptr;
    // the first launch after a harvest (arm) follows ordinary kernels: full serialisation there
    const bool capped = p.stream_cap > 0;
    if (capped) launch_pdl(ns_cn_kernel<DV, DC, true>, gc, dim3(block), st, pdl && !arm && !sample, p);
    else launch_pdl(ns_cn_kernel<DV, DC, false>, gc, dim3(block), st, pdl && !arm && !sample, p);
    if (sample) cudaEventRecord(ev[1], st);
    if (capped) {
        if (arm) launch_pdl(ns_x_kernel<true, true>, gx, dim3(block), st, pdl && !sample, p);
        else launch_pdl(ns_x_kernel<false, true>, gx, dim3(block), st, pdl && !sample, p);
    } else {
        if (arm) launch_pdl(ns_x_kernel<true, false>, gx, dim3(block), st, pdl && !sample, p);
        else launch_pdl(ns_x_kernel<false, false>, gx, dim3(block), st, pdl && !sample, p);
    }
    if (sample) {
        cudaEventRecord(ev[2], st);
        g_prof.iter_idx[g_prof.n_samples++] = p.iter;
    }
}

// one iteration of every graph's frame stream; arm: lanes re-armed by the preceding harvest take their new frames
int bp_launch_node_iteration(int dv, int dc, const BpParams &p, bool arm, cudaStream_t st)
{
    if (dv == 4 && dc == 8) launch_node_iteration<4, 8>(p, arm, st);
    else if (dv == 3 && dc == 6) launch_node_iteration<3, 6>(p, arm, st);
    else if (dv == 5 && dc == 10) launch_node_iteration<5, 10>(p, arm, st);
    else if (dv == 3 && dc == 9) launch_node_iteration<3, 9>(p, arm, st);
    else if (dv == 4 && dc == 12) launch_node_iteration<4, 12>(p, arm, st);
    else return -1;
    return 0;
}

}  // namespace scldpc
