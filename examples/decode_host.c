/* examples/decode_host.c -- calling libscldpc from C exactly where the reference's frame loop calls
 * generate_code(); channel_doped(); decodeBP(); plr_computation();  (BP_FULL.c:2117-2150).
 *
 *   gcc -std=c11 examples/decode_host.c -Iinclude -Lfl_scaling_sc_ldpc_b200 -lscldpc -Wl,-rpath,$PWD/fl_scaling_sc_ldpc_b200 -o decode_host
 *
 * Builds a small (4,8) code with the reference's construction (socket permutations, here from rand()), draws BEC
 * realisations, decodes G x F frames in one call and accumulates the reference's counters.
 * Exit status: 0 = decoded, 2 = library reported an error (without a GPU: "no CUDA device ... no CPU fallback"). */
#include <stdio.h>
#include <stdlib.h>

#include "scldpc.h"

enum { DV = 4, DC = 8, L = 12, CNS = 16, VNS = 32, G = 2, F = 8 };

int main(void)
{
    const int n = L * VNS, len = CNS * DC, npos = L + DV - 1;
    int32_t *vn_cn = malloc(sizeof(int32_t) * G * n * DV);
    uint8_t *erased = malloc((size_t)G * F * n);
    int *perm = malloc(sizeof(int) * len), *inter = malloc(sizeof(int) * npos * len);
    srand(7);
    for (int g = 0; g < G; g++) {
        for (int i = 0; i < len; i++) perm[i] = i;
        for (int p = 0; p < npos; p++) {                              /* generate_code, BP_FULL.c:1679-1699 */
            for (int i = 0; i < len; i++) {
                int k = i + rand() % (len - i), t = perm[i];
                perm[i] = perm[k]; perm[k] = t;
            }
            for (int i = 0; i < len; i++) inter[p * len + i] = p * CNS + perm[i] / DC;
        }
        for (int p = 0; p < L; p++)                                   /* BP_FULL.c:1702-1716 */
            for (int t = 0; t < VNS; t++)
                for (int i = 0; i < DV; i++)
                    vn_cn[((size_t)g * n + p * VNS + t) * DV + i] = inter[(p + i) * len + DV * t + i];
        for (int f = 0; f < F; f++)                                   /* channel, BP_FULL.c:1525-1545 */
            for (int v = 0; v < n; v++) erased[((size_t)g * F + f) * n + v] = ((double)rand() / RAND_MAX) < 0.45;
    }
    scldpc_dims_t d = { DV, DC, L, VNS, CNS, G, /*n_words=*/2, /*n_frames=*/F };
    int32_t iters[G * F], residual[G * F], blocks[G * F], er_exp[G * F], bl_exp[G * F];
    int rc = scldpc_decode_host(&d, vn_cn, erased, /*W=*/0, /*max_it=*/100, 0, SCLDPC_F_TERMINATED, iters, residual, blocks, er_exp,
                                bl_exp, NULL, NULL, NULL, 0);
    if (rc) {
        fprintf(stderr, "libscldpc error %d: %s\n", rc, scldpc_last_error());
        return 2;
    }
    long users_err = 0, frame_err = 0, block_err = 0;                 /* plr_computation, BP_FULL.c:1503-1520 */
    for (int k = 0; k < G * F; k++)
        if (residual[k] > 0) { users_err += residual[k]; frame_err++; block_err += blocks[k]; }
    printf("frames %d  frame_err %ld  users_err %ld  block_err %ld  iterations of frame 0: %d\n", G * F, frame_err, users_err, block_err,
           iters[0]);
    free(vn_cn); free(erased); free(perm); free(inter);
    return 0;
}
