// graph_kernels.cu -- ensemble generation, CN tables, channel realisations (sm_100a).
//
//  * graph tables: vn_cn -> cn_edge / vn_slot in the reference's CN order (generate_code, BP_FULL.c:1702-1716)
//  * on-device "Olmos random ensemble" (generate_code BP_FULL.c:1656-1761 == SC.gen_slots SC.py:33-56): one uniform
//    socket permutation per CN position, drawn by sorting Philox4x32-10 keys (a uniform permutation, unlike a
//    keyed bijection) with a shared-memory + global bitonic network
//  * BEC realisations, bit-sliced, with hard / soft doping (channel_doped BP_FULL.c:1547-1574; PD.py:154,174-192)
#include "common.cuh"

namespace scldpc {

// ------------------------------------------------------------------------------------------------------------
// CN tables
// ------------------------------------------------------------------------------------------------------------
__global__ void graph_fill_kernel(const int32_t *vn_cn, int32_t *cn_edge, int32_t *fill, int n, int nk, int dv, int dc,
                                  int *err)
{
    const int g = blockIdx.y;
    const long long E = (long long)n * dv;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
        const int c = vn_cn[(size_t)g * E + e];
        if (c < 0 || c >= nk) { atomicExch(err, 1); continue; }
        const int j = atomicAdd(fill + (size_t)g * nk + c, 1);
        if (j >= dc) { atomicExch(err, 2); continue; }
        cn_edge[((size_t)g * nk + c) * dc + j] = (int)e;
    }
}

// sort each CN row ascending (= VN order), pad with E, and write the reverse index
__global__ void graph_rows_kernel(int32_t *cn_edge, int32_t *vn_slot, const int32_t *fill, int n, int nk, int dv, int dc)
{
    const int g = blockIdx.y;
    const int E = n * dv;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nk; c += gridDim.x * blockDim.x) {
        int32_t *row = cn_edge + ((size_t)g * nk + c) * dc;
        int deg = fill[(size_t)g * nk + c];
        if (deg > dc) deg = dc;
        for (int j = deg; j < dc; j++) row[j] = E;
        for (int a = 1; a < deg; a++) {          // insertion sort, deg <= dc
            int key = row[a], b = a - 1;
            while (b >= 0 && row[b] > key) { row[b + 1] = row[b]; b--; }
            row[b + 1] = key;
        }
        for (int j = 0; j < deg; j++) vn_slot[(size_t)g * E + row[j]] = c * dc + j;
    }
}

int graph_build_tables(const int32_t *vn_cn, int32_t *vn_slot, int32_t *cn_edge, int32_t *scratch, int *err_dev, int G,
                       int n, int nk, int dv, int dc, cudaStream_t st)
{
    cudaMemsetAsync(scratch, 0, sizeof(int32_t) * (size_t)G * nk, st);
    cudaMemsetAsync(err_dev, 0, sizeof(int), st);
    long long E = (long long)n * dv;
    unsigned bx = (unsigned)((E + 255) / 256 < 1184 ? (E + 255) / 256 : 1184);
    graph_fill_kernel<<<dim3(bx, G), 256, 0, st>>>(vn_cn, cn_edge, scratch, n, nk, dv, dc, err_dev);
    unsigned bc = (unsigned)((nk + 255) / 256 < 1184 ? (nk + 255) / 256 : 1184);
    graph_rows_kernel<<<dim3(bc, G), 256, 0, st>>>(cn_edge, vn_slot, scratch, n, nk, dv, dc);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// ensemble generation
// ------------------------------------------------------------------------------------------------------------
#define SORT_TILE 4096          // keys sorted per block in shared memory (32 KB)
#define SORT_THREADS 512
#define KEY_IDX_BITS 20          // low bits carry the socket index (S <= 2^20)

// keys[seg][s] = (random << KEY_IDX_BITS) | s for s < S, all-ones padding up to Spad.  seg = g*npos + pos.
// pos0: absolute index of the chain's first position (streams decoded in consecutive, overlapping pieces draw the same
// permutation for the same absolute CN position)
__global__ void graph_keys_kernel(u64 *keys, int S, int Spad, int npos, uint64_t seed, uint64_t first_graph, uint32_t pos0)
{
    const int seg = blockIdx.x;                       // segments on grid.x: G * npos can exceed the 65535 limit of grid.y
    const uint64_t gid = first_graph + (uint64_t)(seg / npos);
    const uint32_t pos = pos0 + (uint32_t)(seg % npos);
    for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < Spad / 2; q += gridDim.y * blockDim.x) {
        uint32_t r[4];
        // counter = (socket pair, position, graph id); key = seed ^ domain tag "graph"
        philox4x32_10((uint32_t)q, pos, (uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)seed ^ 0x67726170u,
                      (uint32_t)(seed >> 32), r);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int s = 2 * q + h;
            const u64 rnd = ((u64)r[2 * h] << 32 | r[2 * h + 1]) >> KEY_IDX_BITS;   // 44 random bits
            keys[(size_t)seg * Spad + s] = (s < S) ? ((rnd << KEY_IDX_BITS) | (u64)s) : ~0ull;
        }
    }
}

__device__ __forceinline__ void cmpx(u64 &a, u64 &b, bool up)
{
    if ((a > b) == up) { u64 t = a; a = b; b = t; }
}

// Runs, inside one SORT_TILE tile, all compare-exchange passes of bitonic stages [k_lo, k_hi] whose distance fits the
// tile: for stage k the passes j = min(k-1, log2(tile)-1) .. 0.  (Stage k merges runs of 2^k keys.)
__global__ void __launch_bounds__(SORT_THREADS) bitonic_tile_kernel(u64 *keys, int Spad, int k_lo, int k_hi)
{
    __shared__ u64 s[SORT_TILE];
    const int tile = Spad < SORT_TILE ? Spad : SORT_TILE;
    const size_t base = (size_t)blockIdx.x * Spad + (size_t)blockIdx.y * tile;
    for (int i = threadIdx.x; i < tile; i += SORT_THREADS) s[i] = keys[base + i];
    __syncthreads();
    int lt = 0;
    while ((1 << lt) < tile) lt++;
    for (int k = k_lo; k <= k_hi; k++) {
        for (int j = (k - 1 < lt - 1 ? k - 1 : lt - 1); j >= 0; j--) {
            for (int t = threadIdx.x; t < tile / 2; t += SORT_THREADS) {
                const int lo = ((t >> j) << (j + 1)) | (t & ((1 << j) - 1));
                const int hi = lo | (1 << j);
                const size_t gi = (size_t)blockIdx.y * tile + lo;      // index inside the segment
                const bool up = ((gi >> k) & 1) == 0;
                cmpx(s[lo], s[hi], up);
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < tile; i += SORT_THREADS) keys[base + i] = s[i];
}

// one global compare-exchange pass (distance 2^j >= SORT_TILE) of stage k
__global__ void bitonic_global_kernel(u64 *keys, int Spad, int k, int j)
{
    u64 *seg = keys + (size_t)blockIdx.x * Spad;
    for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < Spad / 2; t += gridDim.y * blockDim.x) {
        const int lo = ((t >> j) << (j + 1)) | (t & ((1 << j) - 1));
        const int hi = lo | (1 << j);
        const bool up = ((lo >> k) & 1) == 0;
        u64 a = seg[lo], b = seg[hi];
        if ((a > b) == up) { seg[lo] = b; seg[hi] = a; }
    }
}

// vn_cn[g][pos*vns_pos+t][i] = cp*cns_pos + perm_cp[dv*t+i]/dc with cp = pos+i (mod L if tail-biting), perm_cp[s] the
// socket index carried by the s-th smallest key of CN position cp  (BP_FULL.c:1693,1712; SC.py:26-38)
__global__ void graph_from_keys_kernel(const u64 *keys, int32_t *vn_cn, int Spad, int npos, int L, int vns_pos, int cns_pos,
                                       int dv, int dc, int tail_biting)
{
    const int g = blockIdx.y;
    const long long E = (long long)L * vns_pos * dv;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % dv);
        const long long v = e / dv;
        const int pos = (int)(v / vns_pos), t = (int)(v % vns_pos);
        int cp = pos + i;
        if (tail_biting) cp %= L;
        const u64 key = keys[((size_t)g * npos + cp) * Spad + dv * t + i];
        const int sock = (int)(key & ((1ull << KEY_IDX_BITS) - 1));
        vn_cn[(size_t)g * E + e] = cp * cns_pos + sock / dc;
    }
}

// Protograph-based ensemble (sc_ldpc_protograph.py:6-20, used at PD.py:198-239): the M VNs of a position form M/cns_pos
// portions of cns_pos VNs; edge i of VN t of a portion goes to CN perm[t] of position p+i, with an independent uniform
// permutation of the cns_pos CNs per (position, portion, edge type).  keys holds one sorted segment per permutation.
__global__ void graph_from_keys_proto_kernel(const u64 *keys, int32_t *vn_cn, int Spad, int L, int vns_pos, int cns_pos, int dv)
{
    const int g = blockIdx.y;
    const int portions = vns_pos / cns_pos;
    const long long E = (long long)L * vns_pos * dv;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % dv);
        const long long v = e / dv;
        const int pos = (int)(v / vns_pos), tt = (int)(v % vns_pos);
        const int portion = tt / cns_pos, t = tt % cns_pos;
        const size_t seg = (((size_t)g * L + pos) * portions + portion) * dv + i;
        const u64 key = keys[seg * Spad + t];
        vn_cn[(size_t)g * E + e] = (pos + i) * cns_pos + (int)(key & ((1ull << KEY_IDX_BITS) - 1));
    }
}

// ensemble: 0 = semi-structured (generate_code / SC.gen_slots), 1 = its tail-biting variant, 2 = protograph-based
int graph_generate(int32_t *vn_cn, u64 *keys, int G, int L, int vns_pos, int cns_pos, int dv, int dc, uint64_t seed,
                   uint64_t first_graph, int ensemble, cudaStream_t st, uint32_t first_position)
{
    if (first_position && ensemble != 0) return -1;
    const bool proto = ensemble == 2;
    const int tail_biting = ensemble == 1;
    if (proto && (vns_pos % cns_pos)) return -1;
    const int S = proto ? cns_pos : cns_pos * dc;
    if (S > (1 << KEY_IDX_BITS)) return -1;
    int lg = 1;
    while ((1 << lg) < S) lg++;
    const int Spad = 1 << lg;
    const int npos = proto ? L * (vns_pos / cns_pos) * dv : (tail_biting ? L : L + dv - 1);   // segments per graph
    const int segs = G * npos;
    unsigned bx = (unsigned)((Spad / 2 + 255) / 256);
    if (bx > 64) bx = 64;
    graph_keys_kernel<<<dim3(segs, bx), 256, 0, st>>>(keys, S, Spad, npos, seed, first_graph, first_position);
    const int tile = Spad < SORT_TILE ? Spad : SORT_TILE;
    int lt = 0;
    while ((1 << lt) < tile) lt++;
    const dim3 gt(segs, Spad / tile);
    // stages 1..lt entirely inside a tile
    bitonic_tile_kernel<<<gt, SORT_THREADS, 0, st>>>(keys, Spad, 1, lt);
    for (int k = lt + 1; k <= lg; k++) {
        for (int j = k - 1; j >= lt; j--) bitonic_global_kernel<<<dim3(segs, bx), 256, 0, st>>>(keys, Spad, k, j);
        bitonic_tile_kernel<<<gt, SORT_THREADS, 0, st>>>(keys, Spad, k, k);
    }
    const long long E = (long long)L * vns_pos * dv;
    unsigned be = (unsigned)((E + 255) / 256 < 1184 ? (E + 255) / 256 : 1184);
    if (proto) graph_from_keys_proto_kernel<<<dim3(be, G), 256, 0, st>>>(keys, vn_cn, Spad, L, vns_pos, cns_pos, dv);
    else graph_from_keys_kernel<<<dim3(be, G), 256, 0, st>>>(keys, vn_cn, Spad, npos, L, vns_pos, cns_pos, dv, dc, tail_biting);
    return 0;
}

size_t graph_generate_scratch_words(int G, int L, int vns_pos, int cns_pos, int dv, int dc, int ensemble)
{
    const bool proto = ensemble == 2;
    const int S = proto ? cns_pos : cns_pos * dc;
    int lg = 1;
    while ((1 << lg) < S) lg++;
    const size_t segs = proto ? (size_t)L * (vns_pos / (cns_pos > 0 ? cns_pos : 1)) * dv : (size_t)(ensemble == 1 ? L : L + dv - 1);
    return (size_t)G * segs * ((size_t)1 << lg);
}

// ------------------------------------------------------------------------------------------------------------
// channel
// ------------------------------------------------------------------------------------------------------------
// chan[g][v][w] bit b = 1 (erased) iff channel_draw(seed; graph, frame first_frame+64w+b, v) < eps * 2^32, unless v is among the first
// known[pos] VNs of its position (doping).  One Philox call yields the draws of 4 consecutive frames.
// v0: absolute index of the chain's first VN (stream pieces, see graph_keys_kernel)
__global__ void channel_generate_kernel(u64 *chan, int n, int W, int n_frames, int vns_pos, const int32_t *known,
                                        u64 thr, uint64_t seed, uint64_t first_graph, uint32_t first_frame, uint32_t v0)
{
    const int g = blockIdx.y;
    const uint64_t gid = first_graph + (uint64_t)g;
    const long long items = (long long)n * W;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < items; idx += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(idx / W), w = (int)(idx % W);
        u64 word = 0;
        const bool forced_known = known && (v % vns_pos) < known[v / vns_pos];
        if (!forced_known) {
#pragma unroll 4
            for (int q = 0; q < 16; q++) {
                uint32_t r[4];
                // frame ids of this Philox call: first_frame + 64w + 4q + {0,1,2,3}; first_frame is a multiple of 4
                philox4x32_10(v0 + (uint32_t)v, (first_frame >> 2) + (uint32_t)(w * 16 + q), (uint32_t)gid, (uint32_t)(gid >> 32),
                              (uint32_t)seed ^ 0x6368616Eu, (uint32_t)(seed >> 32), r);
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const int f = w * 64 + q * 4 + h;
                    if (f < n_frames && (u64)r[h] < thr) word |= 1ull << (q * 4 + h);
                }
            }
        }
        chan[(size_t)g * items + idx] = word;
    }
}

void channel_generate(u64 *chan, int G, int n, int W, int n_frames, int vns_pos, const int32_t *known_dev, double eps,
                      uint64_t seed, uint64_t first_graph, uint32_t first_frame, cudaStream_t st, uint32_t first_vn)
{
    u64 thr;
    if (eps <= 0.0) thr = 0;
    else if (eps >= 1.0) thr = 1ull << 32;
    else thr = (u64)(eps * 4294967296.0);
    const long long items = (long long)n * W;
    unsigned bx = (unsigned)((items + 255) / 256 < 2368 ? (items + 255) / 256 : 2368);
    channel_generate_kernel<<<dim3(bx, G), 256, 0, st>>>(chan, n, W, n_frames, vns_pos, known_dev, thr, seed, first_graph, first_frame, first_vn);
}

// bytes [G][F][n] (1 = erased) -> bit-sliced words [G][n][W]
__global__ void channel_pack_kernel(const uint8_t *bytes, u64 *chan, int n, int W, int F)
{
    const int g = blockIdx.z, w = blockIdx.y;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        u64 word = 0;
        for (int b = 0; b < 64; b++) {
            const int f = w * 64 + b;
            if (f < F && bytes[((size_t)g * F + f) * n + v]) word |= 1ull << b;
        }
        chan[((size_t)g * n + v) * W + w] = word;
    }
}

// bit-sliced words [G][n][W] -> bytes [G][F][n]
__global__ void bits_unpack_kernel(const u64 *bits, uint8_t *bytes, int n, int W, int F)
{
    const int g = blockIdx.z, f = blockIdx.y;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x)
        bytes[((size_t)g * F + f) * n + v] = (uint8_t)((bits[((size_t)g * n + v) * W + (f >> 6)] >> (f & 63)) & 1ull);
}

void channel_pack(const uint8_t *bytes_dev, u64 *chan, int G, int n, int W, int F, cudaStream_t st)
{
    unsigned bx = (unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
    channel_pack_kernel<<<dim3(bx, W, G), 256, 0, st>>>(bytes_dev, chan, n, W, F);
}

void bits_unpack(const u64 *bits, uint8_t *bytes_dev, int G, int n, int W, int F, cudaStream_t st)
{
    if (F <= 0) return;
    unsigned bx = (unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
    bits_unpack_kernel<<<dim3(bx, F, G), 256, 0, st>>>(bits, bytes_dev, n, W, F);
}

}  // namespace scldpc
