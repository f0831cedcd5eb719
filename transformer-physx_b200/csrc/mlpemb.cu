// Lorenz / Rossler Koopman embedding kernels (K3, K4) and the fused clip + Adam step.
// Reference: trphysx/embedding/embedding_lorenz.py:94-105 (embed), :107-118 (recover), :120-142
// (koopmanOperation), :144-157 (koopmanOperator); enn_trainer.py:97-99 (clip_grad_norm_ + optimizer.step).
#include "mlpemb.cuh"
#include "mlpemb_umma.cuh"
#include <cmath>
#include <algorithm>

using namespace trpx;

namespace {

constexpr int EMB_THREADS = 128;
constexpr int SPT = 2;             // samples per thread (weights broadcast from smem are reused SPT times)

// ---- weight packer ---------------------------------------------------------------------------------
struct PackSrc {
    const float *mu, *sd, *W1, *b1, *W2, *b2, *lnw, *lnb, *V1, *c1, *V2, *c2, *kd, *ku;
};

__global__ void mlp_pack_kernel(PackSrc s, float* blob, MlpLayout L) {
    const int Hd = L.Hd;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.total; i += stride) {
        float v = 0.f;
        if (i < L.encB) {                       // encA [Hd][4]
            const int u = (i - L.encA) >> 2, j = (i - L.encA) & 3;
            v = j < 3 ? s.W1[u * MS + j] : s.b1[u];
        } else if (i < L.b2) {                  // encB [Hd][32] = W2[e][u]
            const int u = (i - L.encB) / ME, e = (i - L.encB) % ME;
            v = s.W2[e * Hd + u];
        } else if (i < L.lnw) v = s.b2[i - L.b2];
        else if (i < L.lnb) v = s.lnw[i - L.lnw];
        else if (i < L.decA) v = s.lnb[i - L.lnb];
        else if (i < L.decB) v = s.V1[i - L.decA];
        else if (i < L.c2) {                    // decB [Hd][4]
            const int u = (i - L.decB) >> 2, j = (i - L.decB) & 3;
            v = j < 3 ? s.V2[j * Hd + u] : s.c1[u];
        } else if (i < L.mu) { const int j = i - L.c2; v = j < 3 ? s.c2[j] : 0.f; }
        else if (i < L.sd) { const int j = i - L.mu; v = j < 3 ? s.mu[j] : 0.f; }
        else if (i < L.kd) { const int j = i - L.sd; v = j < 3 ? s.sd[j] : 1.f; }
        else if (i < L.ku) v = s.kd[i - L.kd];
        else if (i < L.W2) { const int j = i - L.ku; v = j < 61 ? s.ku[j] : 0.f; }
        else if (i < L.V2) v = s.W2[i - L.W2];
        else v = s.V2[i - L.V2];
        blob[i] = v;
    }
}

// ---- K3a: embed  x[N,3] -> g[N,32] --------------------------------------------------------------------
// Thread-per-sample (SPT samples each): the 500 hidden units are streamed once; per unit the thread does
// 3 FMAs (layer 1), a ReLU and 32 FMAs into its 32 output accumulators with weights BROADCAST from smem
// (9 LDS.128 per unit shared by SPT samples).  LayerNorm(32) finishes in registers.
// Optionally saves the normalised pre-affine activation and 1/std for the training backward.
__global__ void __launch_bounds__(EMB_THREADS) mlp_embed_kernel(const float* __restrict__ blob, MlpLayout L,
                                                                const float* __restrict__ x, float* __restrict__ g,
                                                                long long N, float eps, float* __restrict__ zhat_out,
                                                                float* __restrict__ rstd_out) {
    extern __shared__ float4 sm4[];
    const int Hd = L.Hd;
    float4* sA = sm4;                 // [Hd]
    float4* sB = sm4 + Hd;            // [Hd][8]
    {
        const float4* src = reinterpret_cast<const float4*>(blob + L.encA);
        for (int i = threadIdx.x; i < Hd * 9; i += blockDim.x) sm4[i] = __ldg(src + i);
    }
    __syncthreads();
    float mu[MS], sd[MS];
#pragma unroll
    for (int j = 0; j < MS; j++) { mu[j] = __ldg(blob + L.mu + j); sd[j] = __ldg(blob + L.sd + j); }
    const long long step = (long long)gridDim.x * blockDim.x * SPT;
    for (long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * SPT; base < N; base += step) {
        float xn[SPT][MS];
#pragma unroll
        for (int s = 0; s < SPT; s++)
#pragma unroll
            for (int j = 0; j < MS; j++) {
                const float v = (base + s < N) ? x[(base + s) * MS + j] : mu[j];
                xn[s][j] = (v - mu[j]) / sd[j];           // _normalize, embedding_lorenz.py:159-160
            }
        float acc[SPT][ME];
#pragma unroll
        for (int e = 0; e < ME; e++) {
            const float b = __ldg(blob + L.b2 + e);
#pragma unroll
            for (int s = 0; s < SPT; s++) acc[s][e] = b;
        }
#pragma unroll 2
        for (int u = 0; u < Hd; u++) {
            const float4 a = sA[u];
            float hid[SPT];
#pragma unroll
            for (int s = 0; s < SPT; s++)
                hid[s] = fmaxf(fmaf(a.z, xn[s][2], fmaf(a.y, xn[s][1], fmaf(a.x, xn[s][0], a.w))), 0.f);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 w = sB[u * 8 + j];
                // (acc[e], acc[e+1]) += (w[e], w[e+1]) * hid: one FFMA2 per weight pair (scalar-broadcast multiplier)
#pragma unroll
                for (int s = 0; s < SPT; s++) {
                    ffma2(acc[s][4 * j + 0], acc[s][4 * j + 1], w.x, w.y, hid[s]);
                    ffma2(acc[s][4 * j + 2], acc[s][4 * j + 3], w.z, w.w, hid[s]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < SPT; s++) {
            if (base + s >= N) break;
            float m = 0.f;
#pragma unroll
            for (int e = 0; e < ME; e++) m += acc[s][e];
            m *= (1.0f / ME);
            float v = 0.f;
#pragma unroll
            for (int e = 0; e < ME; e++) {
                acc[s][e] -= m;
                v = fmaf(acc[s][e], acc[s][e], v);
            }
            const float rstd = 1.0f / sqrtf(v * (1.0f / ME) + eps);
            float4* go = reinterpret_cast<float4*>(g + (base + s) * ME);
            float4* zo = zhat_out ? reinterpret_cast<float4*>(zhat_out + (base + s) * ME) : nullptr;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                float4 z, o;
                z.x = acc[s][4 * j + 0] * rstd; z.y = acc[s][4 * j + 1] * rstd;
                z.z = acc[s][4 * j + 2] * rstd; z.w = acc[s][4 * j + 3] * rstd;
                o.x = z.x * __ldg(blob + L.lnw + 4 * j + 0) + __ldg(blob + L.lnb + 4 * j + 0);
                o.y = z.y * __ldg(blob + L.lnw + 4 * j + 1) + __ldg(blob + L.lnb + 4 * j + 1);
                o.z = z.z * __ldg(blob + L.lnw + 4 * j + 2) + __ldg(blob + L.lnb + 4 * j + 2);
                o.w = z.w * __ldg(blob + L.lnw + 4 * j + 3) + __ldg(blob + L.lnb + 4 * j + 3);
                go[j] = o;
                if (zo) zo[j] = z;
            }
            if (rstd_out) rstd_out[base + s] = rstd;
        }
    }
}

// ---- K3a (small N): the same map for batches that cannot fill the machine with one thread per sample -------------------
// (the Koopman training step embeds 8,192 samples, a 256-trajectory rollout 256: the thread-per-sample kernel then runs
// 16 CTAs / 1 CTA, and its 500-unit loop is a 78 us latency chain).  Here a WARP owns 8 samples: the lanes first share
// the 500 hidden units (16 each) and leave relu(W1 x + b1) in shared memory, then lane = output channel streams the units
// once (two broadcast LDS.128 for the 8 hidden values, one conflict-free word of W2^T, 8 FMAs); LayerNorm over the lanes.
constexpr int EMS_S = 8;               // samples per warp
constexpr int EMS_WARPS = 8;
__global__ void __launch_bounds__(EMS_WARPS * 32) mlp_embed_small_kernel(const float* __restrict__ blob, MlpLayout L,
                                                                          const float* __restrict__ x, float* __restrict__ g,
                                                                          long long N, float eps, float* __restrict__ zhat_out,
                                                                          float* __restrict__ rstd_out) {
    extern __shared__ float4 sm4[];
    const int Hd = L.Hd, HP = (Hd + 31) & ~31;
    float4* sA = sm4;                                   // [Hd]     {W1[u][0..2], b1[u]}
    const float* sB = reinterpret_cast<const float*>(sm4 + Hd);      // [Hd][32] W2^T
    float4* hb = sm4 + Hd * 9 + (threadIdx.x >> 5) * (HP * 2);        // [HP][8]  hidden values of the warp's samples
    {
        const float4* src = reinterpret_cast<const float4*>(blob + L.encA);
        for (int i = threadIdx.x; i < Hd * 9; i += blockDim.x) sm4[i] = __ldg(src + i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * EMS_WARPS + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * EMS_WARPS;
    float mu[MS], sd[MS];
#pragma unroll
    for (int j = 0; j < MS; j++) { mu[j] = __ldg(blob + L.mu + j); sd[j] = __ldg(blob + L.sd + j); }
    const float b2 = __ldg(blob + L.b2 + lane), lw = __ldg(blob + L.lnw + lane), lb = __ldg(blob + L.lnb + lane);
    for (long long base = warp * EMS_S; base < N; base += nwarps * EMS_S) {
        float xn[EMS_S][MS];
#pragma unroll
        for (int s = 0; s < EMS_S; s++)
#pragma unroll
            for (int j = 0; j < MS; j++) {
                const float v = (base + s < N) ? __ldg(x + (base + s) * MS + j) : mu[j];
                xn[s][j] = (v - mu[j]) / sd[j];           // _normalize, embedding_lorenz.py:159-160
            }
        for (int u = lane; u < HP; u += 32) {
            float hid[EMS_S];
            if (u < Hd) {
                const float4 a = sA[u];
#pragma unroll
                for (int s = 0; s < EMS_S; s++)
                    hid[s] = fmaxf(fmaf(a.z, xn[s][2], fmaf(a.y, xn[s][1], fmaf(a.x, xn[s][0], a.w))), 0.f);
            } else {
#pragma unroll
                for (int s = 0; s < EMS_S; s++) hid[s] = 0.f;
            }
            hb[u * 2] = make_float4(hid[0], hid[1], hid[2], hid[3]);
            hb[u * 2 + 1] = make_float4(hid[4], hid[5], hid[6], hid[7]);
        }
        __syncwarp();
        float acc[EMS_S];
#pragma unroll
        for (int s = 0; s < EMS_S; s++) acc[s] = b2;
#pragma unroll 4
        for (int u = 0; u < Hd; u++) {
            const float4 h0 = hb[u * 2], h1 = hb[u * 2 + 1];
            const float w = sB[u * ME + lane];
            acc[0] = fmaf(w, h0.x, acc[0]); acc[1] = fmaf(w, h0.y, acc[1]);
            acc[2] = fmaf(w, h0.z, acc[2]); acc[3] = fmaf(w, h0.w, acc[3]);
            acc[4] = fmaf(w, h1.x, acc[4]); acc[5] = fmaf(w, h1.y, acc[5]);
            acc[6] = fmaf(w, h1.z, acc[6]); acc[7] = fmaf(w, h1.w, acc[7]);
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < EMS_S; s++) {
            const float m = warp_sum(acc[s]) * (1.0f / ME);
            const float d = acc[s] - m;
            const float rstd = 1.0f / sqrtf(warp_sum(d * d) * (1.0f / ME) + eps);
            if (base + s < N) {
                const float z = d * rstd;
                g[(base + s) * ME + lane] = z * lw + lb;
                if (zhat_out) zhat_out[(base + s) * ME + lane] = z;
                if (rstd_out && lane == 0) rstd_out[base + s] = rstd;
            }
        }
    }
}

// ---- K3b: recover  g[N,32] -> x[N,3] ------------------------------------------------------------------
// With `target` != null the kernel is the fused `recover + MSE` of Trainer.eval_states (utils/trainer.py:380-390): the
// recovered states are not written (x may be null); each CTA writes the sum of squared errors of its samples to
// partial[blockIdx.x] (summed in a fixed order by sum_partials_kernel: deterministic).
__global__ void __launch_bounds__(EMB_THREADS) mlp_recover_kernel(const float* __restrict__ blob, MlpLayout L,
                                                                  const float* __restrict__ g, float* __restrict__ x,
                                                                  long long N, const float* __restrict__ target,
                                                                  float* __restrict__ partial) {
    __shared__ float red[EMB_THREADS / 32];
    float sq = 0.f;
    extern __shared__ float4 sm4[];
    const int Hd = L.Hd;
    float4* sA = sm4;                 // [Hd][8]  V1 rows
    float4* sB = sm4 + Hd * 8;        // [Hd]     {V2[0..2][u], c1[u]}
    {
        const float4* src = reinterpret_cast<const float4*>(blob + L.decA);
        for (int i = threadIdx.x; i < Hd * 9; i += blockDim.x) sm4[i] = __ldg(src + i);
    }
    __syncthreads();
    const long long step = (long long)gridDim.x * blockDim.x * SPT;
    for (long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * SPT; base < N; base += step) {
        float gi[SPT][ME];
#pragma unroll
        for (int s = 0; s < SPT; s++) {
            const long long n = (base + s < N) ? base + s : N - 1;
            const float4* gp = reinterpret_cast<const float4*>(g + n * ME);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 v = __ldg(gp + j);
                gi[s][4 * j] = v.x; gi[s][4 * j + 1] = v.y; gi[s][4 * j + 2] = v.z; gi[s][4 * j + 3] = v.w;
            }
        }
        float y[SPT][MS];
#pragma unroll
        for (int s = 0; s < SPT; s++)
#pragma unroll
            for (int c = 0; c < MS; c++) y[s][c] = __ldg(blob + L.c2 + c);
#pragma unroll 2
        for (int u = 0; u < Hd; u++) {
            const float4 b = sB[u];
            float p0[SPT], p1[SPT];
#pragma unroll
            for (int s = 0; s < SPT; s++) { p0[s] = b.w; p1[s] = 0.f; }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 w = sA[u * 8 + j];
                // the (even, odd) partial sums of a sample advance together: one FFMA2 per weight pair, operands
                // in the register order the 128-bit loads left them in
#pragma unroll
                for (int s = 0; s < SPT; s++) {
                    ffma2v(p0[s], p1[s], gi[s][4 * j + 0], gi[s][4 * j + 1], w.x, w.y);
                    ffma2v(p0[s], p1[s], gi[s][4 * j + 2], gi[s][4 * j + 3], w.z, w.w);
                }
            }
#pragma unroll
            for (int s = 0; s < SPT; s++) {
                const float hid = fmaxf(p0[s] + p1[s], 0.f);
                ffma2v(y[s][0], y[s][1], hid, hid, b.x, b.y);
                y[s][2] = fmaf(hid, b.z, y[s][2]);
            }
        }
#pragma unroll
        for (int s = 0; s < SPT; s++) {
            if (base + s >= N) break;
#pragma unroll
            for (int c = 0; c < MS; c++) {                   // _unnormalize, embedding_lorenz.py:162-163
                const float v = __ldg(blob + L.sd + c) * y[s][c] + __ldg(blob + L.mu + c);
                if (x != nullptr) x[(base + s) * MS + c] = v;
                if (target != nullptr) {
                    const float d = v - target[(base + s) * MS + c];
                    sq = fmaf(d, d, sq);
                }
            }
        }
    }
    if (partial != nullptr) {
        sq = warp_sum(sq);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < EMB_THREADS / 32; i++) t += red[i];
            partial[blockIdx.x] = t;
        }
    }
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < n; i++) t += partial[i];
        out[0] = t * scale;
    }
}

// ---- K3b (tensor cores): recover for large N --------------------------------------------------------------
// Layer 1 ([N x 32] . [32 x 500], 91 % of the FLOPs of recover) on the warp-level TF32 MMA with the error-compensated
// 3xTF32 product (a_lo w_hi + a_hi w_lo + a_hi w_hi, FP32 accumulate: ~2^-21 relative, inside the FP32 parity bar).
// One warp = one 16-sample M-tile; the weights are split into TF32 hi / lo planes ONCE per CTA into shared memory
// ([Hp][36] each, row stride 36 makes the B-fragment loads conflict-free), the A fragments (the 16 x 32 inputs) are
// split once per tile and reused for all 63 n-tiles.  The accumulator fragment of an n-tile (2 hidden units x 2
// samples per lane) goes through ReLU straight into layer 2 (12 FMAs per lane against one LDS.128 per hidden unit);
// the three outputs are reduced over the 4 lanes of a row at the end.  Instructions per 16 samples: ~2,900 against
// ~5,500 for the SIMT kernel.
constexpr int RTC_THREADS = 384;
constexpr int RTC_LD = 36;

__device__ __forceinline__ uint32_t rtc_tf32(float x) {
    return tf32_rna_bits(x);
}
__device__ __forceinline__ void rtc_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// With `target` != null it is also the fused recover + MSE (x may then be null): per-CTA squared-error sums go to
// partial[blockIdx.x], as in the FP32 kernel.
__global__ void __launch_bounds__(RTC_THREADS, 1) mlp_recover_tc_kernel(const float* __restrict__ blob, MlpLayout L,
                                                                        const float* __restrict__ g,
                                                                        float* __restrict__ x, long long N,
                                                                        const float* __restrict__ target,
                                                                        float* __restrict__ partial) {
    extern __shared__ __align__(16) float smf[];
    __shared__ float red[RTC_THREADS / 32];
    float sq = 0.f;
    const int Hd = L.Hd, Hp = (Hd + 7) & ~7;
    uint32_t* whi = reinterpret_cast<uint32_t*>(smf);         // [Hp][36] TF32 hi plane of V1[n][k]
    uint32_t* wlo = whi + Hp * RTC_LD;                         // lo plane (FP32 remainder)
    float4* dB = reinterpret_cast<float4*>(wlo + Hp * RTC_LD); // [Hp] {V2[0..2][n], c1[n]}
    for (int i = threadIdx.x; i < Hp * ME; i += blockDim.x) {
        const int n = i / ME, k = i - n * ME;
        const float w = n < Hd ? __ldg(blob + L.decA + i) : 0.f;
        const uint32_t hi = rtc_tf32(w);
        whi[n * RTC_LD + k] = hi;
        wlo[n * RTC_LD + k] = __float_as_uint(w - __uint_as_float(hi));
    }
    for (int i = threadIdx.x; i < Hp; i += blockDim.x)
        dB[i] = i < Hd ? __ldg(reinterpret_cast<const float4*>(blob + L.decB) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int gq = lane >> 2, q = lane & 3;
    const long long ntiles = (N + 15) / 16;
    float c2[MS], mu[MS], sd[MS];
#pragma unroll
    for (int c = 0; c < MS; c++) { c2[c] = __ldg(blob + L.c2 + c); mu[c] = __ldg(blob + L.mu + c); sd[c] = __ldg(blob + L.sd + c); }
    for (long long tile = (long long)blockIdx.x * nw + warp; tile < ntiles; tile += (long long)gridDim.x * nw) {
        const long long r0 = tile * 16 + gq, r1 = r0 + 8;
        const float* g0 = g + (r0 < N ? r0 : N - 1) * ME;
        const float* g1 = g + (r1 < N ? r1 : N - 1) * ME;
        uint32_t ah[4][4], al[4][4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const float a[4] = {__ldg(g0 + 8 * s + q), __ldg(g1 + 8 * s + q), __ldg(g0 + 8 * s + q + 4), __ldg(g1 + 8 * s + q + 4)};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                ah[s][i] = rtc_tf32(a[i]);
                al[s][i] = __float_as_uint(a[i] - __uint_as_float(ah[s][i]));
            }
        }
        float y0[MS] = {0.f, 0.f, 0.f}, y1[MS] = {0.f, 0.f, 0.f};
#pragma unroll 2
        for (int t = 0; t < Hp / 8; t++) {
            const float4 v0 = dB[8 * t + 2 * q], v1 = dB[8 * t + 2 * q + 1];
            float acc[4] = {v0.w, v1.w, v0.w, v1.w};          // bias c1 of the lane's two hidden units
            const uint32_t* bh = whi + (8 * t + gq) * RTC_LD + q;
            const uint32_t* bl = wlo + (8 * t + gq) * RTC_LD + q;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const uint32_t bh0 = bh[8 * s], bh1 = bh[8 * s + 4], bl0 = bl[8 * s], bl1 = bl[8 * s + 4];
                rtc_mma(acc, al[s], bh0, bh1);
                rtc_mma(acc, ah[s], bl0, bl1);
                rtc_mma(acc, ah[s], bh0, bh1);
            }
            const float h00 = fmaxf(acc[0], 0.f), h01 = fmaxf(acc[1], 0.f), h10 = fmaxf(acc[2], 0.f), h11 = fmaxf(acc[3], 0.f);
            y0[0] = fmaf(h00, v0.x, y0[0]); y0[1] = fmaf(h00, v0.y, y0[1]); y0[2] = fmaf(h00, v0.z, y0[2]);
            y0[0] = fmaf(h01, v1.x, y0[0]); y0[1] = fmaf(h01, v1.y, y0[1]); y0[2] = fmaf(h01, v1.z, y0[2]);
            y1[0] = fmaf(h10, v0.x, y1[0]); y1[1] = fmaf(h10, v0.y, y1[1]); y1[2] = fmaf(h10, v0.z, y1[2]);
            y1[0] = fmaf(h11, v1.x, y1[0]); y1[1] = fmaf(h11, v1.y, y1[1]); y1[2] = fmaf(h11, v1.z, y1[2]);
        }
#pragma unroll
        for (int c = 0; c < MS; c++) {
            y0[c] += __shfl_xor_sync(FULL, y0[c], 1);
            y0[c] += __shfl_xor_sync(FULL, y0[c], 2);
            y1[c] += __shfl_xor_sync(FULL, y1[c], 1);
            y1[c] += __shfl_xor_sync(FULL, y1[c], 2);
        }
        if (q == 0) {
#pragma unroll
            for (int c = 0; c < MS; c++) {                   // _unnormalize, embedding_lorenz.py:162-163
                const float v0 = sd[c] * (y0[c] + c2[c]) + mu[c], v1 = sd[c] * (y1[c] + c2[c]) + mu[c];
                if (x != nullptr) {
                    if (r0 < N) x[r0 * MS + c] = v0;
                    if (r1 < N) x[r1 * MS + c] = v1;
                }
                if (target != nullptr) {
                    if (r0 < N) { const float d = v0 - target[r0 * MS + c]; sq = fmaf(d, d, sq); }
                    if (r1 < N) { const float d = v1 - target[r1 * MS + c]; sq = fmaf(d, d, sq); }
                }
            }
        }
    }
    if (partial != nullptr) {
        sq = warp_sum(sq);
        if (lane == 0) red[warp] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < nw; i++) t += red[i];
            partial[blockIdx.x] = t;
        }
    }
}

// ---- K4: banded Koopman matvec ------------------------------------------------------------------------
// K = diag(d) + U - U^T with U holding offsets +1 (u1[0..30]) and +2 (u2[0..29])  (embedding_lorenz.py:59-66,
// 130-137).  g'[i] = d[i] g[i] + u1[i] g[i+1] + u2[i] g[i+2] - u1[i-1] g[i-1] - u2[i-2] g[i-2].
__global__ void mlp_koopman_kernel(const float* __restrict__ blob, MlpLayout L, const float* __restrict__ g,
                                   float* __restrict__ out, long long N) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * ME) return;
    const int i = (int)(idx % ME);
    const float* gr = g + (idx - i);
    const float* d = blob + L.kd;
    const float* u1 = blob + L.ku;
    const float* u2 = blob + L.ku + (ME - 1);
    // same left-to-right order as the dense row-times-vector product of torch.bmm
    float acc = 0.f;
    if (i >= 2) acc = fmaf(-__ldg(u2 + i - 2), gr[i - 2], acc);
    if (i >= 1) acc = fmaf(-__ldg(u1 + i - 1), gr[i - 1], acc);
    acc = fmaf(__ldg(d + i), gr[i], acc);
    if (i + 1 < ME) acc = fmaf(__ldg(u1 + i), gr[i + 1], acc);
    if (i + 2 < ME) acc = fmaf(__ldg(u2 + i), gr[i + 2], acc);
    out[idx] = acc;
}

__global__ void mlp_kmatrix_kernel(const float* __restrict__ blob, MlpLayout L, float* __restrict__ K) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ME * ME) return;
    const int i = idx / ME, j = idx % ME;
    const float* u1 = blob + L.ku;
    const float* u2 = blob + L.ku + (ME - 1);
    float v = 0.f;
    if (j == i) v = blob[L.kd + i];
    else if (j == i + 1) v = u1[i];
    else if (j == i + 2) v = u2[i];
    else if (j == i - 1) v = -u1[j];
    else if (j == i - 2) v = -u2[j];
    K[idx] = v;
}

// ---- fused clip_grad_norm_ + Adam ---------------------------------------------------------------------
constexpr int NORM_BLOCKS = 64;

__global__ void sqnorm_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
    __shared__ float red[8];
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s = fmaf(g[i], g[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
        partial[blockIdx.x] = t;
    }
}

struct AdamArgs {
    float lr, beta1, beta2, eps, wd, max_norm, bc1, bc2_sqrt;
};

__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, const float* __restrict__ partial, AdamArgs a,
                                 float* __restrict__ norm_out) {
    // every block reduces the partial sums in the same fixed order -> identical clip coefficient everywhere
    float tot = 0.f;
    for (int i = 0; i < NORM_BLOCKS; i++) tot += partial[i];
    const float norm = sqrtf(tot);
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
    const float coef = fminf(a.max_norm / (norm + 1e-6f), 1.0f);
    const float step_size = a.lr / a.bc1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * coef;
        const float pi = p[i];
        gi = fmaf(a.wd, pi, gi);                             // L2 weight decay folded into the gradient (Adam, not AdamW)
        const float mi = m[i] + (1.0f - a.beta1) * (gi - m[i]);       // exp_avg.lerp_(grad, 1-beta1)
        const float vi = fmaf(1.0f - a.beta2, gi * gi, v[i] * a.beta2);
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / a.bc2_sqrt + a.eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

float* g_norm_partial[16] = {nullptr};      // per-device scratch for trpx_clip_adam

}  // namespace

extern "C" {

int trpx_mlpemb_create(const trpx_mlpemb_config* cfg, int device, trpx_mlpemb_t* out) {
    TRPX_REQUIRE(cfg != nullptr && out != nullptr, "trpx_mlpemb_create: null argument");
    TRPX_REQUIRE(cfg->state_dim == MS, "state_dim=%d unsupported (3)", cfg->state_dim);
    TRPX_REQUIRE(cfg->n_embd == ME, "n_embd=%d unsupported (32)", cfg->n_embd);
    TRPX_REQUIRE(cfg->hidden > 0 && cfg->hidden <= 1024 && cfg->hidden % 4 == 0, "hidden=%d unsupported", cfg->hidden);
    int ndev = 0;
    TRPX_CUDA(cudaGetDeviceCount(&ndev));
    TRPX_REQUIRE(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
    DeviceGuard guard(device);
    trpx_mlpemb* h = new (std::nothrow) trpx_mlpemb();
    TRPX_REQUIRE(h != nullptr, "out of host memory");
    h->cfg = *cfg;
    h->device = device;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    const MlpLayout L = MlpLayout::make(cfg->hidden);
    if (cudaMalloc(&h->blob, sizeof(float) * L.total) != cudaSuccess) {
        delete h;
        return set_error(TRPX_E_CUDA, "cudaMalloc of the weight blob failed");
    }
    *out = h;
    return TRPX_OK;
}

int trpx_mlpemb_destroy(trpx_mlpemb_t h) {
    if (!valid(h)) return set_error(TRPX_E_STATE, "trpx_mlpemb_destroy: invalid handle");
    DeviceGuard guard(h->device);
    cudaFree(h->blob);
    cudaFree(h->ws);
    cudaFree(h->mse_partial);
    h->magic = 0;
    delete h;
    return TRPX_OK;
}

long long trpx_mlpemb_num_params(trpx_mlpemb_t h) {
    if (!valid(h)) return -1;
    const long long Hd = h->cfg.hidden;
    return ME + 61 + Hd * MS + Hd + ME * Hd + ME + ME + ME + Hd * ME + Hd + MS * Hd + MS;
}

int trpx_mlpemb_load_weights(trpx_mlpemb_t h, const float* const* w, int n, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    TRPX_REQUIRE(w != nullptr && n == 14, "expected 14 weight tensors, got %d", n);
    for (int i = 0; i < n; i++) TRPX_REQUIRE(w[i] != nullptr, "weight tensor %d is null", i);
    DeviceGuard guard(h->device);
    PackSrc s{w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8], w[9], w[10], w[11], w[12], w[13]};
    const MlpLayout L = MlpLayout::make(h->cfg.hidden);
    mlp_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(s, h->blob, L);
    TRPX_CUDA(cudaGetLastError());
    h->loaded = true;
    return TRPX_OK;
}

static int emb_grid(trpx_mlpemb_t h, long long N, int ctas_per_sm) {
    const long long need = (N + (long long)EMB_THREADS * SPT - 1) / ((long long)EMB_THREADS * SPT);
    const long long cap = (long long)h->sm_count * ctas_per_sm;
    return (int)(need < cap ? need : cap);
}

}  // extern "C"

namespace trpx {
int mlpemb_embed_ex(trpx_mlpemb_t h, const float* x, float* g, long long N, float* zhat, float* rstd,
                    void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "weights not loaded");
    TRPX_REQUIRE(N >= 0 && (N == 0 || (x != nullptr && g != nullptr)), "bad arguments");
    if (N == 0) return TRPX_OK;
    DeviceGuard guard(h->device);
    const MlpLayout L = MlpLayout::make(h->cfg.hidden);
    const size_t smem = sizeof(float4) * (size_t)L.Hd * 9;
    // batches that give the thread-per-sample kernel less than ~one CTA per SM: the warp-per-8-samples kernel
    const size_t smem_small = smem + sizeof(float4) * (size_t)EMS_WARPS * (((size_t)L.Hd + 31) & ~(size_t)31) * 2;
    if (N <= (long long)h->sm_count * EMB_THREADS * SPT && smem_small + 1024 <= (size_t)h->smem_optin) {
        const long long nwarps = (N + EMS_S - 1) / EMS_S;
        const int grid = (int)std::min<long long>(h->sm_count, (nwarps + EMS_WARPS - 1) / EMS_WARPS);
        TRPX_CUDA(cudaFuncSetAttribute(mlp_embed_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_small));
        mlp_embed_small_kernel<<<grid, EMS_WARPS * 32, smem_small, (cudaStream_t)stream>>>(h->blob, L, x, g, N, h->cfg.ln_eps,
                                                                                           zhat, rstd);
        TRPX_CUDA(cudaGetLastError());
        return TRPX_OK;
    }
    TRPX_CUDA(cudaFuncSetAttribute(mlp_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_embed_kernel<<<emb_grid(h, N, 3), EMB_THREADS, smem, (cudaStream_t)stream>>>(h->blob, L, x, g, N, h->cfg.ln_eps,
                                                                                    zhat, rstd);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

}  // namespace trpx

extern "C" {

int trpx_mlpemb_embed(trpx_mlpemb_t h, const float* x, float* g, long long N, void* stream) {
    return trpx::mlpemb_embed_ex(h, x, g, N, nullptr, nullptr, stream);
}

int trpx_mlpemb_recover(trpx_mlpemb_t h, const float* g, float* x, long long N, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "weights not loaded");
    TRPX_REQUIRE(N >= 0 && (N == 0 || (x != nullptr && g != nullptr)), "bad arguments");
    if (N == 0) return TRPX_OK;
    DeviceGuard guard(h->device);
    const MlpLayout L = MlpLayout::make(h->cfg.hidden);
    const size_t smem = sizeof(float4) * (size_t)L.Hd * 9;
    TRPX_CUDA(cudaFuncSetAttribute(mlp_recover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int Hp = (L.Hd + 7) & ~7;
    const size_t smem_tc = sizeof(float) * ((size_t)2 * Hp * RTC_LD + (size_t)4 * Hp);
    if ((h->recover_mode == 0 || h->recover_mode == 3) && mlp_recover_umma_ok(h, N))
        return mlp_recover_umma(h, g, x, N, nullptr, nullptr, (cudaStream_t)stream);      // tcgen05 kernel (mlpemb_umma.cu)
    if (h->recover_mode != 1 && N >= 16384 && smem_tc + 1024 <= (size_t)h->smem_optin) {
        // tensor-core kernel: one persistent CTA per SM, 12 warps, 16 samples per warp-tile
        TRPX_CUDA(cudaFuncSetAttribute(mlp_recover_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
        const long long ntiles = (N + 15) / 16;
        const int grid = (int)std::min<long long>(h->sm_count, (ntiles + RTC_THREADS / 32 - 1) / (RTC_THREADS / 32));
        mlp_recover_tc_kernel<<<grid, RTC_THREADS, smem_tc, (cudaStream_t)stream>>>(h->blob, L, g, x, N, nullptr, nullptr);
        TRPX_CUDA(cudaGetLastError());
        return TRPX_OK;
    }
    mlp_recover_kernel<<<emb_grid(h, N, 3), EMB_THREADS, smem, (cudaStream_t)stream>>>(h->blob, L, g, x, N, nullptr,
                                                                                        nullptr);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

int trpx_mlpemb_set_recover_mode(trpx_mlpemb_t h, int mode) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    TRPX_REQUIRE(mode >= 0 && mode <= 3, "mode must be 0 (automatic), 1 (FP32 SIMT kernel only), 2 (no tcgen05 kernel) or 3 (same as 0)");
    h->recover_mode = mode;
    return TRPX_OK;
}

int trpx_mlpemb_recover_mse(trpx_mlpemb_t h, const float* g, const float* target, long long N, float* states_or_null,
                            float* mse_dev, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "weights not loaded");
    TRPX_REQUIRE(N >= 1 && g != nullptr && target != nullptr && mse_dev != nullptr, "bad arguments");
    DeviceGuard guard(h->device);
    const MlpLayout L = MlpLayout::make(h->cfg.hidden);
    const size_t smem = sizeof(float4) * (size_t)L.Hd * 9;
    const int Hp = (L.Hd + 7) & ~7;
    const size_t smem_tc = sizeof(float) * ((size_t)2 * Hp * RTC_LD + (size_t)4 * Hp);
    const bool umma = (h->recover_mode == 0 || h->recover_mode == 3) && mlp_recover_umma_ok(h, N);
    const bool tc = h->recover_mode != 1 && N >= 16384 && smem_tc + 1024 <= (size_t)h->smem_optin;
    const long long ntiles = (N + 15) / 16;
    const int grid = umma ? mlp_recover_umma_grid(h, N)
                     : tc ? (int)std::min<long long>(h->sm_count, (ntiles + RTC_THREADS / 32 - 1) / (RTC_THREADS / 32))
                          : emb_grid(h, N, 3);
    if (h->mse_partial == nullptr || h->mse_partial_n < grid) {
        if (h->mse_partial) TRPX_CUDA(cudaFree(h->mse_partial));
        h->mse_partial = nullptr;
        TRPX_CUDA(cudaMalloc(&h->mse_partial, sizeof(float) * grid));
        h->mse_partial_n = grid;
    }
    if (umma) {
        int rc = mlp_recover_umma(h, g, states_or_null, N, target, h->mse_partial, (cudaStream_t)stream);
        if (rc) return rc;
    } else if (tc) {
        TRPX_CUDA(cudaFuncSetAttribute(mlp_recover_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
        mlp_recover_tc_kernel<<<grid, RTC_THREADS, smem_tc, (cudaStream_t)stream>>>(h->blob, L, g, states_or_null, N, target,
                                                                                     h->mse_partial);
    } else {
        TRPX_CUDA(cudaFuncSetAttribute(mlp_recover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_recover_kernel<<<grid, EMB_THREADS, smem, (cudaStream_t)stream>>>(h->blob, L, g, states_or_null, N, target,
                                                                               h->mse_partial);
    }
    TRPX_CUDA(cudaGetLastError());
    sum_partials_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->mse_partial, grid, 1.0f / ((float)N * (float)MS), mse_dev);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

int trpx_mlpemb_koopman(trpx_mlpemb_t h, const float* g, float* gnext, long long N, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "weights not loaded");
    TRPX_REQUIRE(N >= 0 && (N == 0 || (g != nullptr && gnext != nullptr)), "bad arguments");
    TRPX_REQUIRE(g != gnext, "in-place Koopman step is not supported");
    if (N == 0) return TRPX_OK;
    DeviceGuard guard(h->device);
    const MlpLayout L = MlpLayout::make(h->cfg.hidden);
    const long long total = N * ME;
    mlp_koopman_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->blob, L, g, gnext, N);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

int trpx_mlpemb_koopman_matrix(trpx_mlpemb_t h, float* K, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "weights not loaded");
    TRPX_REQUIRE(K != nullptr, "null output");
    DeviceGuard guard(h->device);
    const MlpLayout L = MlpLayout::make(h->cfg.hidden);
    mlp_kmatrix_kernel<<<(ME * ME + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->blob, L, K);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

int trpx_clip_adam(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float beta1,
                   float beta2, float eps, float weight_decay, float max_norm, float* norm_out, int device,
                   void* stream) {
    TRPX_REQUIRE(p && g && m && v && n > 0, "bad arguments");
    TRPX_REQUIRE(step >= 1, "step must be >= 1");
    TRPX_REQUIRE(device >= 0 && device < 16, "device %d out of range", device);
    DeviceGuard guard(device);
    if (g_norm_partial[device] == nullptr) TRPX_CUDA(cudaMalloc(&g_norm_partial[device], sizeof(float) * NORM_BLOCKS));
    cudaStream_t st = (cudaStream_t)stream;
    sqnorm_partial_kernel<<<NORM_BLOCKS, 256, 0, st>>>(g, n, g_norm_partial[device]);
    TRPX_CUDA(cudaGetLastError());
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay; a.max_norm = max_norm;
    a.bc1 = (float)(1.0 - std::pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)std::sqrt(1.0 - std::pow((double)beta2, (double)step));
    const int blocks = (int)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
    clip_adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, g_norm_partial[device], a, norm_out);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

}  // extern "C"
