// peel_kernels.cu -- random-order peeling decoder with the per-step degree-one-CN trajectory (sm_100a).
//
// Replaces the hot loop of simulate_peeling_decoder_ldpc (PD.py:705-789): at every step pick, uniformly at random, one
// of the check nodes of residual degree one "in ascending CN order" (pick_random_deg_1_cn, PD.py:1022-1026:
// np.flatnonzero(r == 1) + random.choice), remove its variable node from the residual graph, and record how many
// degree-one CNs are left.  The reference pays O(#CN) twice per step (flatnonzero, count_nonzero); here
//   * one warp owns one frame; frames are independent, so thousands run concurrently and hide each other's latency;
//   * the degree-one set is a bitmap in shared memory under a 3-level popcount hierarchy (32 x 32 x 32 fan-out),
//     so "the k-th degree-one CN in ascending order" is three warp-wide prefix sums + one __fns;
//   * CN state lives in global memory as one 64-bit word per CN: (residual degree << 32) + (sum of the ids of the
//     erased VNs still attached).  A CN of degree one therefore names its VN directly (head(schedule[m]), PD.py:769),
//     and removing a VN is one 64-bit atomicAdd per edge whose return value says whether the CN became / stopped
//     being degree one.
// The random pick of step s of frame f is k = philox4x32_10(counter = (s/4, 0, f.lo, f.hi), key = seed)[s%4] mod
// (number of degree-one CNs); the oracle is fed the same 32-bit draws, so trajectories are compared bit for bit.
#include <cstdlib>

#include "common.cuh"

namespace scldpc {

__host__ __device__ __forceinline__ void philox_peel(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                     uint32_t k1, uint32_t (&out)[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// host helper (tests / reproducibility): the 32-bit draws of one frame
void peel_picks_host(uint64_t seed, uint64_t frame_id, int n, uint32_t *out)
{
    for (int s = 0; s < n; s += 4) {
        uint32_t r[4];
        philox_peel((uint32_t)(s >> 2), 0u, (uint32_t)frame_id, (uint32_t)(frame_id >> 32), (uint32_t)seed ^ 0x7065656Cu,
                    (uint32_t)(seed >> 32), r);
        for (int h = 0; h < 4 && s + h < n; h++) out[s + h] = r[h];
    }
}


// position of the (k+1)-th set bit of w (0 <= k < popc(w)): binary search on popcounts (__fns is a bit-by-bit loop in libdevice)
__device__ __forceinline__ int select_bit(unsigned w, int k)
{
    int r = 0, t;
    t = __popc(w & 0xffffu); if (k >= t) { k -= t; r += 16; w >>= 16; }
    t = __popc(w & 0xffu);   if (k >= t) { k -= t; r += 8;  w >>= 8; }
    t = __popc(w & 0xfu);    if (k >= t) { k -= t; r += 4;  w >>= 4; }
    t = __popc(w & 0x3u);    if (k >= t) { k -= t; r += 2;  w >>= 2; }
    t = (int)(w & 1u);       if (k >= t) r += 1;
    return r;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// One warp per frame, PEEL_WPB independent warps per block (a step is a chain of three dependent global round trips -- CN
// state, adjacency row, degree updates -- so throughput is warps in flight / latency: 64 warps per SM instead of the 32 that
// one-warp blocks allow).  Dynamic shared memory per warp: [bitmap words |] level-1 counts | level-2 counts.  When a
// warp's bitmap does not leave room for 64 frames per SM it lives in global memory behind the warp's CN state (it is touched
// by a handful of words per step); the counts always stay in shared memory.
// BITS_GLOBAL is a template parameter so that the bitmap's address space is static: through a generic pointer the bit updates
// compiled to generic ATOM instead of ATOMS / RED.
#define PEEL_WPB 2
// Register targets (measured, profiles/r02z_peel_register_ab.txt): with the bitmap in shared memory (M = 1000) 47 registers / 42
// warps per SM are best (the kernel is close to issue-bound there); with the bitmap in global memory (M >= 10000) the chain of HBM
// round trips wants every warp the SM can hold: 32 registers / 64 warps, +10 %.
template <bool BITS_GLOBAL>
__global__ void __launch_bounds__(32 * PEEL_WPB, BITS_GLOBAL ? 28 : 20) peel_trajectory_kernel(PeelParams p)
{
    extern __shared__ unsigned s_mem_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * PEEL_WPB + warp, n_slots = gridDim.x * PEEL_WPB;
    unsigned *s_mem = s_mem_all + (size_t)warp * ((BITS_GLOBAL ? 0 : p.n_words1) + p.n_l1 + p.n_l2);
    const size_t st_words = (size_t)p.n_cn_all + (BITS_GLOBAL ? ((size_t)p.n_words1 + 1) / 2 : 0);   // u64 per warp
    u64 *st = p.state + (size_t)slot * st_words;
    unsigned *bits_g = reinterpret_cast<unsigned *>(st + p.n_cn_all);   // BITS_GLOBAL
    unsigned *bits_s = s_mem;                                           // !BITS_GLOBAL
    int *l1 = reinterpret_cast<int *>(s_mem + (BITS_GLOBAL ? 0 : p.n_words1));
    int *l2 = l1 + p.n_l1;
    const int l2q = (p.n_l2 + 31) / 32;                           // level-2 entries per lane
    const long long total_frames = (long long)p.G * p.n_frames;

    // the first frame of a warp is its slot number; further frames come from a counter, so warps whose frames stall early (most do
    // at eps = 0.48 without termination) take over work from the ones that peel to the end
    for (long long fr = slot; fr < total_frames;) {
        const int g = (int)(fr / p.n_frames), f = (int)(fr % p.n_frames);
        const int32_t *vn_cn = p.vn_cn + (size_t)g * p.n * p.dv;
        const u64 *chan = p.chan + (size_t)g * p.n * p.W + (f >> 6);
        const int fb = f & 63;
        // ---- residual graph of the erased VNs (schedule, PD.py:750-757) ----
        for (int i = lane; i < p.n_cn_all; i += 32) st[i] = 0;
        for (int i = lane; i < p.n_words1; i += 32) { if (BITS_GLOBAL) bits_g[i] = 0; else bits_s[i] = 0; }
        for (int i = lane; i < p.n_l1 + p.n_l2; i += 32) l1[i] = 0;
        __syncwarp();
        int n_er = 0;
        // (keeping several channel words in flight per lane here was measured slower: the registers cost resident warps)
        for (int v0 = 0; v0 < p.n; v0 += 32) {
            const int v = v0 + lane;
            if (v < p.n && ((chan[(size_t)v * p.W] >> fb) & 1ull)) {
                n_er++;
                for (int i = 0; i < p.dv; i++)
                    atomicAdd(reinterpret_cast<unsigned long long *>(st + vn_cn[(size_t)v * p.dv + i]), (1ull << 32) + (u64)v);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) n_er += __shfl_xor_sync(0xffffffffu, n_er, o);
        __threadfence_block();
        __syncwarp();
        // ---- degree-one bitmap over the CNs the decoder sees (t < total_size, PD.py:756-758) ----
        int cnt1 = 0;
        for (int w = 0; w < p.n_words1; w++) {
            const int c = w * 32 + lane;
            const bool one = c < p.total_size && (__ldcg(reinterpret_cast<const unsigned long long *>(st + c)) >> 32) == 1ull;
            const unsigned m = __ballot_sync(0xffffffffu, one);
            if (lane == 0 && m) {
                if (BITS_GLOBAL) bits_g[w] = m; else bits_s[w] = m;
                const int pc = __popc(m);
                l1[w >> 5] += pc;
                l2[w >> 10] += pc;
            }
            cnt1 += __popc(m);
        }
        __syncwarp();
        int32_t *r1 = p.r1 ? p.r1 + (size_t)fr * (p.num_steps + 1) : nullptr;
        if (r1 && lane == 0) r1[0] = cnt1;
        const uint64_t fid = p.first_frame + (uint64_t)fr;
        int recovered = 0;
        uint32_t rnd[4] = {0, 0, 0, 0};
        int step = 0;
        for (; step < p.num_steps; step++) {
            if (cnt1 == 0) break;                               // stalled or finished: r1 stays put (PD.py:765-767)
            if ((step & 3) == 0)
                philox_peel((uint32_t)(step >> 2), 0u, (uint32_t)fid, (uint32_t)(fid >> 32), (uint32_t)p.seed ^ 0x7065656Cu,
                            (uint32_t)(p.seed >> 32), rnd);
            const int h = step & 3;                             // selects instead of a dynamic index: rnd stays in registers
            const uint32_t draw = h == 0 ? rnd[0] : (h == 1 ? rnd[1] : (h == 2 ? rnd[2] : rnd[3]));
            int k = (int)(draw % (uint32_t)cnt1);
            // ---- select the k-th set bit ----
            int a = 0;
            for (int q = 0; q < l2q; q++) a += (lane * l2q + q) < p.n_l2 ? l2[lane * l2q + q] : 0;
            int inc = warp_incl_scan(a, lane);
            unsigned bal = __ballot_sync(0xffffffffu, k < inc);
            int sel = __ffs(bal) - 1;
            k -= __shfl_sync(0xffffffffu, inc - a, sel);
            int b2 = sel * l2q;
            for (int q = 0; q + 1 < l2q; q++) {                    // at most a few entries inside the selected lane
                const int c2 = l2[b2];
                if (k < c2) break;
                k -= c2;
                b2++;
            }
            a = (b2 * 32 + lane) < p.n_l1 ? l1[b2 * 32 + lane] : 0;
            inc = warp_incl_scan(a, lane);
            bal = __ballot_sync(0xffffffffu, k < inc);
            sel = __ffs(bal) - 1;
            k -= __shfl_sync(0xffffffffu, inc - a, sel);
            const int b1 = b2 * 32 + sel;
            const int wi = b1 * 32 + lane;
            const unsigned word = wi < p.n_words1 ? (BITS_GLOBAL ? __ldcg(bits_g + wi) : bits_s[wi]) : 0u;
            a = __popc(word);
            inc = warp_incl_scan(a, lane);
            bal = __ballot_sync(0xffffffffu, k < inc);
            sel = __ffs(bal) - 1;
            k -= __shfl_sync(0xffffffffu, inc - a, sel);
            const unsigned wsel = __shfl_sync(0xffffffffu, word, sel);
            const int m = (b1 * 32 + sel) * 32 + select_bit(wsel, k);
            // ---- remove its VN (PD.py:769-780) ----
            const int v = (int)(uint32_t)__ldcg(reinterpret_cast<const unsigned long long *>(st + m));
            recovered++;
            int delta = 0;
            if (lane < p.dv) {
                const int c = vn_cn[(size_t)v * p.dv + lane];
                const u64 old = atomicAdd(reinterpret_cast<unsigned long long *>(st + c), ~((1ull << 32) + (u64)v) + 1ull);
                const int deg_old = (int)(old >> 32);
                if (c < p.total_size) {                          // CNs >= total_size are not seen by the decoder
                    if (deg_old == 2) {                          // becomes degree one
                        if (BITS_GLOBAL) atomicOr(&bits_g[c >> 5], 1u << (c & 31)); else atomicOr(&bits_s[c >> 5], 1u << (c & 31));
                        atomicAdd(&l1[c >> 10], 1);
                        atomicAdd(&l2[c >> 15], 1);
                        delta = 1;
                    } else if (deg_old == 1) {                   // stops being degree one
                        if (BITS_GLOBAL) atomicAnd(&bits_g[c >> 5], ~(1u << (c & 31))); else atomicAnd(&bits_s[c >> 5], ~(1u << (c & 31)));
                        atomicSub(&l1[c >> 10], 1);
                        atomicSub(&l2[c >> 15], 1);
                        delta = -1;
                    }
                }
            }
            cnt1 += __reduce_add_sync(0xffffffffu, delta);
            __syncwarp();
            if (r1 && lane == 0) r1[step + 1] = cnt1;             // PD.py:781
        }
        if (r1)
            for (int s = step + 1 + lane; s <= p.num_steps; s += 32) r1[s] = cnt1;   // copies of the last value
        if (lane == 0) {
            p.recovered[fr] = recovered;
            p.n_erased[fr] = n_er;
            fr = (long long)n_slots + (long long)atomicAdd(reinterpret_cast<unsigned long long *>(p.next), 1ull);
        }
        fr = __shfl_sync(0xffffffffu, fr, 0);
    }
}

// calc_var_chunk (EST.py:131-138) on the device: for step s, over the frames in order,
//   ssq[s] += (r1/M - theory[s]/M)^2 for frames with r1 != 0,   counts[s] += (r1 != 0)
// Sums run over frames sequentially in double precision without FMA contraction, the order np.nansum(axis=0) uses.
__global__ void peel_variance_kernel(const int32_t *r1, int n_frames, int row_len, const double *theory, int S, double M,
                                     double *ssq, long long *counts)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const double th = __ddiv_rn(theory[s], M);
    double acc = 0.0;
    long long cnt = 0;
    for (int f = 0; f < n_frames; f++) {
        const int v = r1[(size_t)f * row_len + s];
        if (v != 0) {
            const double c = __dsub_rn(__ddiv_rn((double)v, M), th);
            acc = __dadd_rn(acc, __dmul_rn(c, c));
            cnt++;
        }
    }
    ssq[s] = __dadd_rn(ssq[s], acc);
    counts[s] += cnt;
}

// shared memory per block; *bits_global = 1 when the bitmap has to live in global memory
size_t peel_smem_bytes(int total_size, int *n_words1, int *n_l1, int *n_l2, int *bits_global)
{
    const int w = (total_size + 31) / 32, a = (w + 31) / 32, b = (a + 31) / 32;
    if (n_words1) *n_words1 = w;
    if (n_l1) *n_l1 = a;
    if (n_l2) *n_l2 = b;
    size_t limit = 3400;                                                     // 64 frames per SM in flight (227 KB / 64)
    if (const char *e = getenv("SCLDPC_PEEL_SMEM_LIMIT")) limit = (size_t)atol(e);   // tests force the global-bitmap path
    const bool big = sizeof(unsigned) * ((size_t)w + a + b) > limit;
    if (bits_global) *bits_global = big ? 1 : 0;
    return sizeof(unsigned) * ((big ? 0 : (size_t)w) + a + b);
}

// u64 words of workspace per block
size_t peel_state_words(int n_cn_all, int total_size)
{
    int w = 0, big = 0;
    peel_smem_bytes(total_size, &w, nullptr, nullptr, &big);
    return (size_t)n_cn_all + (big ? ((size_t)w + 1) / 2 : 0);
}

// number of frames decoded concurrently (= warps of the launch = CN-state slots of the workspace)
int peel_grid(int total_size, long long total_frames, int n_cn_all)
{
    int dev = 0, sms = 148, max_smem = 227 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int n_l2 = 0;
    const size_t smem = peel_smem_bytes(total_size, nullptr, nullptr, &n_l2, nullptr);
    if (smem > (size_t)max_smem || n_l2 > 32 * 8) return -1;
    // frames in flight per SM = the warps that are really resident (registers allow fewer than the 64 the shared memory is sized
    // for): a grid beyond that runs as a second, mostly empty wave
    int big = 0, occ = 0;
    peel_smem_bytes(total_size, nullptr, nullptr, nullptr, &big);
    auto kernel = big ? peel_trajectory_kernel<true> : peel_trajectory_kernel<false>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem * PEEL_WPB));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 32 * PEEL_WPB, smem * PEEL_WPB) != cudaSuccess || occ < 1) occ = 1;
    long long per_sm = (long long)occ * PEEL_WPB;
    long long grid = per_sm * sms;
    // keep the CN-state workspace (8 B per CN and frame in flight) under 24 GB: M = 1e5 holds 21 MB per frame
    const long long by_mem = (24ll << 30) / (8ll * (long long)peel_state_words(n_cn_all, total_size));
    if (grid > by_mem) grid = by_mem;
    if (grid > total_frames) grid = total_frames;
    if (const char *e = getenv("SCLDPC_PEEL_SLOTS")) { const long long cap = atoll(e); if (cap > 0 && grid > cap) grid = cap; }   // timing experiments
    grid = (grid + PEEL_WPB - 1) / PEEL_WPB * PEEL_WPB;          // whole blocks
    return (int)(grid < PEEL_WPB ? PEEL_WPB : grid);
}

int peel_launch(PeelParams p, int grid, cudaStream_t st)
{
    const size_t smem = peel_smem_bytes(p.total_size, &p.n_words1, &p.n_l1, &p.n_l2, &p.bits_global);
    if (p.n_l2 > 32 * 8) return -1;                              // more than 2^23 CNs: one more level would be needed
    auto kernel = p.bits_global ? peel_trajectory_kernel<true> : peel_trajectory_kernel<false>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem * PEEL_WPB)) != cudaSuccess) return -2;
    g_prof.launches += 1;
    kernel<<<grid / PEEL_WPB, 32 * PEEL_WPB, smem * PEEL_WPB, st>>>(p);
    return 0;
}

void peel_variance_launch(const int32_t *r1, int n_frames, int row_len, const double *theory, int S, double M, double *ssq,
                          long long *counts, cudaStream_t st)
{
    g_prof.launches += 1;
    peel_variance_kernel<<<(S + 255) / 256, 256, 0, st>>>(r1, n_frames, row_len, theory, S, M, ssq, counts);
}

}  // namespace scldpc
