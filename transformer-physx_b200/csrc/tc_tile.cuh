// Tile products on tensor cores for CTA-cooperative kernels (8 warps): error-compensated 3xTF32 mma.sync.m16n8k8 with both
// operands addressed through (row stride, column stride), so transposed operands need no staging.  Used by the transformer
// training step (gpt2_train_tc.cu) and the Koopman training step (mlpemb_train.cu).
#pragma once
#include "common.cuh"

namespace trpx {
namespace tct {

constexpr int TT = 256;                    // threads per CTA (8 warps)

__device__ __forceinline__ void split3(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rna_bits(x);
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_t(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// C[b][M x N] = A[b][M x K] * B[b][K x N] on tensor cores (3xTF32), b < nb:
//   A element (m, k) = A[b*abs + m*ars + k*acs],  B element (k, n) = B[b*bbs + k*brs + n*bcs]   (shared or global memory)
//   M = 16 * mtiles, N = 8 * ntiles, K = 8 * KS; units are dealt round-robin over the 8 warps.
//   epi(b, m, n, value) receives every element once.  All 256 threads must call; no synchronisation inside.
//   Register blocking: a unit = (batch, group of MT m-tiles, group of NTG n-tiles); per k-step the A fragments of the MT
//   m-tiles and the B fragments of the NTG n-tiles are loaded and split ONCE and feed MT*NTG*3 MMAs (the first version,
//   one n-tile per unit, spent 18 shared-memory loads and 54 split instructions per 12 MMAs and was MIO / issue bound).
//   Fragment slot q of a k-step holds logical k = 2q (a0, a1, b0) and 2q+1 (a2, a3, b1): any permutation of k inside a
//   k-step is a valid contraction order, and this one makes the two k of a slot ADJACENT, so an operand whose k runs along
//   memory (AV: acs == 1, BV: brs == 1) is fetched with one 64-bit load per pair.
//   B fragments are requested PF k-steps ahead (weights come from L2: shared memory leaves almost no L1).
template <int MT, int NTG, int KS, bool AV, bool BV, class Epi>
__device__ __forceinline__ void tc_gemm(const float* A, int abs_, int ars, int acs, const float* B, int bbs, int brs, int bcs,
                                        int nb, int mtiles, int ntiles, Epi epi, int wrot = 0) {
    // wrot rotates the warp -> unit assignment so that consecutive small calls of one phase land on different warps
    const int warp = ((threadIdx.x >> 5) + 8 - wrot) & 7, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    constexpr int PF = KS < 3 ? KS : 3;
    const int mgroups = mtiles / MT, ngroups = ntiles / NTG;
    const int units = nb * mgroups * ngroups;
    for (int u = warp; u < units; u += TT / 32) {
        const int ng = u % ngroups, r = u / ngroups, mg = r % mgroups, b = r / mgroups;
        const float* Ab = A + (size_t)b * abs_ + (size_t)(mg * MT * 16 + g) * ars + (size_t)(2 * q) * acs;
        const float* Bb = B + (size_t)b * bbs + (size_t)(2 * q) * brs + (size_t)(ng * NTG * 8 + g) * bcs;
        float bw[PF][NTG][2];
        auto load_b = [&](int ks, float (&dst)[NTG][2]) {
#pragma unroll
            for (int j = 0; j < NTG; j++) {
                const float* bp = Bb + (size_t)(8 * ks) * brs + (size_t)(8 * j) * bcs;
                if constexpr (BV) {
                    const float2 v = *reinterpret_cast<const float2*>(bp);
                    dst[j][0] = v.x; dst[j][1] = v.y;
                } else {
                    dst[j][0] = bp[0]; dst[j][1] = bp[brs];
                }
            }
        };
#pragma unroll
        for (int ks = 0; ks < PF; ks++) load_b(ks, bw[ks]);
        float acc[MT][NTG][4];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NTG; j++) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            uint32_t ah[MT][4], al[MT][4];
#pragma unroll
            for (int i = 0; i < MT; i++) {
                const float* ap = Ab + (size_t)(16 * i) * ars + (size_t)(8 * ks) * acs;
                float a0, a1, a2, a3;
                if constexpr (AV) {
                    const float2 lo2 = *reinterpret_cast<const float2*>(ap), hi2 = *reinterpret_cast<const float2*>(ap + 8 * ars);
                    a0 = lo2.x; a2 = lo2.y; a1 = hi2.x; a3 = hi2.y;
                } else {
                    a0 = ap[0]; a1 = ap[8 * ars]; a2 = ap[acs]; a3 = ap[8 * ars + acs];
                }
                split3(a0, ah[i][0], al[i][0]);
                split3(a1, ah[i][1], al[i][1]);
                split3(a2, ah[i][2], al[i][2]);
                split3(a3, ah[i][3], al[i][3]);
            }
#pragma unroll
            for (int j = 0; j < NTG; j++) {
                uint32_t bh0, bl0, bh1, bl1;
                split3(bw[ks % PF][j][0], bh0, bl0);
                split3(bw[ks % PF][j][1], bh1, bl1);
#pragma unroll
                for (int i = 0; i < MT; i++) {
                    mma_t(acc[i][j], al[i], bh0, bh1);
                    mma_t(acc[i][j], ah[i], bl0, bl1);
                    mma_t(acc[i][j], ah[i], bh0, bh1);
                }
            }
            if (ks + PF < KS) load_b(ks + PF, bw[ks % PF]);
        }
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NTG; j++) {
                const int m = (mg * MT + i) * 16 + g, n = (ng * NTG + j) * 8 + 2 * q;
                epi(b, m, n, acc[i][j][0]);
                epi(b, m, n + 1, acc[i][j][1]);
                epi(b, m + 8, n, acc[i][j][2]);
                epi(b, m + 8, n + 1, acc[i][j][3]);
            }
    }
}

}  // namespace tct
}  // namespace trpx
