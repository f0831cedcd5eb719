// K7: fused forward + backward of the Lorenz/Rossler Koopman embedding loss.
// Reference: LorenzEmbeddingTrainer.forward, trphysx/embedding/embedding_lorenz.py:181-221 (restated in
// SURVEY.md Appendix D), followed by loss.backward() (embedding/training/enn_trainer.py:93-96).
//
//   enc(x) = LN(W2 relu(W1 (x-mu)/sd + b1) + b2) ,  dec(g) = sd (V2 relu(V1 g + c1) + c2) + mu
//   loss = sum_t w_recon MSE(dec(enc(x_t)), x_t) + sum_{t>=1} [ w_dyn MSE(dec(K^t enc(x_0)), x_t) + w_reg |K|_F^2 ]
//
// Pipeline (all on one stream, deterministic reductions, no atomics):
//   1. mlp_embed_kernel           g, zhat, rstd for the N = B*T encoder samples            (mlpemb.cu)
//   2. chain_fwd_kernel           gk_t = K gk_{t-1} for the B*(T-1) Koopman-path rows
//   3. dec_fwd_bwd_kernel         decoder forward, loss terms, decoder weight grads, d/dg   (M = N + B(T-1) rows)
//   4. chain_bwd_kernel           back through the Koopman recurrence: dK bands, d/dg_0
//   5. enc_bwd_kernel             LayerNorm + encoder backward, encoder weight grads
//   6. finalize_kernel            ordered sum of per-CTA partials -> flat gradient (parameters() order), losses
#include "mlpemb.cuh"
#include "tc_tile.cuh"
#include <algorithm>
#include <cstdlib>

using namespace trpx;

namespace {

constexpr int TR_THREADS = 256;
constexpr int TILE = 32;            // samples per tile
constexpr int NU = 2;               // hidden units owned per thread (Hd <= NU * TR_THREADS)

struct TrainDims {
    int B, T, N, M, Hd;
    float w_recon, w_dyn, w_reg, inv3B;
};

// ---- 2. Koopman chain forward: one warp per trajectory, lane = observable index -----------------------
__device__ __forceinline__ float koopman_apply(float g, float d, float u1, float u2, float u1m, float u2m, int lane) {
    // same term order as mlp_koopman_kernel (j ascending)
    const float gm2 = __shfl_up_sync(FULL, g, 2), gm1 = __shfl_up_sync(FULL, g, 1);
    const float gp1 = __shfl_down_sync(FULL, g, 1), gp2 = __shfl_down_sync(FULL, g, 2);
    float acc = 0.f;
    if (lane >= 2) acc = fmaf(-u2m, gm2, acc);
    if (lane >= 1) acc = fmaf(-u1m, gm1, acc);
    acc = fmaf(d, g, acc);
    if (lane + 1 < ME) acc = fmaf(u1, gp1, acc);
    if (lane + 2 < ME) acc = fmaf(u2, gp2, acc);
    return acc;
}

struct BandRegs {
    float d, u1, u2, u1m, u2m;      // d[i], u1[i], u2[i], u1[i-1], u2[i-2]
};
__device__ __forceinline__ BandRegs load_band(const float* blob, const MlpLayout& L, int lane) {
    BandRegs r;
    const float* u1 = blob + L.ku;
    const float* u2 = blob + L.ku + (ME - 1);
    r.d = blob[L.kd + lane];
    r.u1 = lane < ME - 1 ? u1[lane] : 0.f;
    r.u2 = lane < ME - 2 ? u2[lane] : 0.f;
    r.u1m = lane >= 1 ? u1[lane - 1] : 0.f;
    r.u2m = lane >= 2 ? u2[lane - 2] : 0.f;
    return r;
}

__global__ void chain_fwd_kernel(const float* __restrict__ blob, MlpLayout L, float* __restrict__ G, TrainDims D) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const BandRegs k = load_band(blob, L, lane);
    for (int b = warp; b < D.B; b += nwarps) {
        float g = G[((size_t)b * D.T) * ME + lane];
        for (int t = 1; t < D.T; t++) {
            g = koopman_apply(g, k.d, k.u1, k.u2, k.u1m, k.u2m, lane);
            G[((size_t)D.N + (size_t)b * (D.T - 1) + (t - 1)) * ME + lane] = g;
        }
    }
}

// ---- 3. decoder forward + backward over all M rows --------------------------------------------------------
// per-CTA partial layout
__host__ __device__ inline int dec_part_size(int Hd) { return Hd * ME + Hd + MS * Hd + 4 + 4; }
__host__ __device__ inline int enc_part_size(int Hd) { return Hd * MS + Hd + ME * Hd + 3 * ME; }

__global__ void __launch_bounds__(TR_THREADS, 1) dec_fwd_bwd_kernel(const float* __restrict__ blob, MlpLayout L,
                                                                    const float* __restrict__ states,
                                                                    const float* __restrict__ G, float* __restrict__ dG,
                                                                    float* __restrict__ part, TrainDims D) {
    extern __shared__ __align__(16) float sm[];
    const int Hd = D.Hd;
    float* g_s = sm;                       // [TILE][32]
    float* h_s = g_s + TILE * ME;          // [TILE][Hd]
    float* dy_s = h_s + TILE * Hd;         // [TILE][4]
    float* red = dy_s + TILE * 4;          // [8 warps][8]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* V1 = blob + L.decA;       // [Hd][32]
    const float* dB = blob + L.decB;       // [Hd][4] = {V2[0][u], V2[1][u], V2[2][u], c1[u]}

    float v1r[NU][ME], dv1[NU][ME], v2r[NU][MS], dv2[NU][MS], c1r[NU], dc1[NU];
#pragma unroll
    for (int j = 0; j < NU; j++) {
        const int u = tid + j * TR_THREADS;
        const bool ok = u < Hd;
#pragma unroll
        for (int k = 0; k < ME; k++) {
            v1r[j][k] = ok ? V1[u * ME + k] : 0.f;
            dv1[j][k] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < MS; c++) {
            v2r[j][c] = ok ? dB[u * 4 + c] : 0.f;
            dv2[j][c] = 0.f;
        }
        c1r[j] = ok ? dB[u * 4 + 3] : 0.f;
        dc1[j] = 0.f;
    }
    float w_dc2[MS] = {0.f, 0.f, 0.f}, w_loss = 0.f, w_rec = 0.f;     // lane-0 accumulators of each warp
    float mu[MS], sd[MS], c2[MS];
#pragma unroll
    for (int c = 0; c < MS; c++) { mu[c] = blob[L.mu + c]; sd[c] = blob[L.sd + c]; c2[c] = blob[L.c2 + c]; }

    const int ntiles = (D.M + TILE - 1) / TILE;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int r0 = tile * TILE;
        for (int i = tid; i < TILE * ME; i += TR_THREADS) {
            const int r = r0 + i / ME;
            g_s[i] = r < D.M ? G[(size_t)r * ME + (i % ME)] : 0.f;
        }
        __syncthreads();
        // phase 1: hidden activations of the owned units for every sample of the tile
#pragma unroll
        for (int j = 0; j < NU; j++) {
            const int u = tid + j * TR_THREADS;
            if (u < Hd) {
                for (int s = 0; s < TILE; s++) {
                    const float4* gr = reinterpret_cast<const float4*>(g_s + s * ME);
                    float p0 = c1r[j], p1 = 0.f;
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 gv = gr[q];
                        p0 = fmaf(v1r[j][4 * q + 0], gv.x, p0);
                        p1 = fmaf(v1r[j][4 * q + 1], gv.y, p1);
                        p0 = fmaf(v1r[j][4 * q + 2], gv.z, p0);
                        p1 = fmaf(v1r[j][4 * q + 3], gv.w, p1);
                    }
                    h_s[s * Hd + u] = fmaxf(p0 + p1, 0.f);
                }
            }
        }
        __syncthreads();
        // phase 2: outputs, loss terms and d(loss)/dy for 4 samples per warp
        for (int q = 0; q < TILE / 8; q++) {
            const int s = warp * (TILE / 8) + q;
            const int r = r0 + s;
            float y[MS] = {0.f, 0.f, 0.f};
            for (int u = lane; u < Hd; u += 32) {
                const float hv = h_s[s * Hd + u];
                const float4 w = *reinterpret_cast<const float4*>(dB + u * 4);
                y[0] = fmaf(hv, w.x, y[0]);
                y[1] = fmaf(hv, w.y, y[1]);
                y[2] = fmaf(hv, w.z, y[2]);
            }
#pragma unroll
            for (int c = 0; c < MS; c++) y[c] = warp_sum(y[c]);
            if (lane == 0) {
                float dyv[MS] = {0.f, 0.f, 0.f};
                if (r < D.M) {
                    // row -> (b, t) of its target state; rows >= N belong to the Koopman path
                    int b, t;
                    float a;
                    if (r < D.N) { b = r / D.T; t = r - b * D.T; a = D.w_recon; }
                    else { const int rr = r - D.N; b = rr / (D.T - 1); t = rr - b * (D.T - 1) + 1; a = D.w_dyn; }
                    const float* xt = states + ((size_t)b * D.T + t) * MS;
                    float sq = 0.f;
#pragma unroll
                    for (int c = 0; c < MS; c++) {
                        const float xh = sd[c] * (y[c] + c2[c]) + mu[c];
                        const float df = xh - xt[c];
                        sq = fmaf(df, df, sq);
                        dyv[c] = sd[c] * (a * 2.0f * df * D.inv3B);
                        w_dc2[c] += dyv[c];
                    }
                    w_loss += a * sq * D.inv3B;
                    if (r < D.N) w_rec += sq * D.inv3B;
                }
                dy_s[s * 4 + 0] = dyv[0]; dy_s[s * 4 + 1] = dyv[1]; dy_s[s * 4 + 2] = dyv[2]; dy_s[s * 4 + 3] = 0.f;
            }
        }
        __syncthreads();
        // phase 3: per-unit backward; h_s is overwritten in place with d(loss)/d(hidden pre-activation)
#pragma unroll
        for (int j = 0; j < NU; j++) {
            const int u = tid + j * TR_THREADS;
            if (u < Hd) {
                for (int s = 0; s < TILE; s++) {
                    const float hv = h_s[s * Hd + u];
                    const float4 dyv = *reinterpret_cast<const float4*>(dy_s + s * 4);
                    dv2[j][0] = fmaf(dyv.x, hv, dv2[j][0]);
                    dv2[j][1] = fmaf(dyv.y, hv, dv2[j][1]);
                    dv2[j][2] = fmaf(dyv.z, hv, dv2[j][2]);
                    float dh = fmaf(v2r[j][2], dyv.z, fmaf(v2r[j][1], dyv.y, v2r[j][0] * dyv.x));
                    dh = hv > 0.f ? dh : 0.f;
                    dc1[j] += dh;
                    h_s[s * Hd + u] = dh;
                    const float4* gr = reinterpret_cast<const float4*>(g_s + s * ME);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 gv = gr[q];
                        dv1[j][4 * q + 0] = fmaf(dh, gv.x, dv1[j][4 * q + 0]);
                        dv1[j][4 * q + 1] = fmaf(dh, gv.y, dv1[j][4 * q + 1]);
                        dv1[j][4 * q + 2] = fmaf(dh, gv.z, dv1[j][4 * q + 2]);
                        dv1[j][4 * q + 3] = fmaf(dh, gv.w, dv1[j][4 * q + 3]);
                    }
                }
            }
        }
        __syncthreads();
        // phase 4: d(loss)/dg[s][k] = sum_u V1[u][k] dh[s][u] ; thread = (sample, 4 consecutive k)
        {
            const int s = tid >> 3, k4 = (tid & 7) * 4;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int u = 0; u < Hd; u++) {
                const float dh = h_s[s * Hd + u];
                const float4 w = __ldg(reinterpret_cast<const float4*>(V1 + u * ME + k4));
                a0 = fmaf(dh, w.x, a0); a1 = fmaf(dh, w.y, a1); a2 = fmaf(dh, w.z, a2); a3 = fmaf(dh, w.w, a3);
            }
            const int r = r0 + s;
            if (r < D.M) *reinterpret_cast<float4*>(dG + (size_t)r * ME + k4) = make_float4(a0, a1, a2, a3);
        }
        __syncthreads();
    }
    // per-CTA partials
    float* P = part + (size_t)blockIdx.x * dec_part_size(Hd);
#pragma unroll
    for (int j = 0; j < NU; j++) {
        const int u = tid + j * TR_THREADS;
        if (u < Hd) {
#pragma unroll
            for (int k = 0; k < ME; k++) P[u * ME + k] = dv1[j][k];
            P[Hd * ME + u] = dc1[j];
#pragma unroll
            for (int c = 0; c < MS; c++) P[Hd * ME + Hd + c * Hd + u] = dv2[j][c];
        }
    }
    if (lane == 0) {
        red[warp * 8 + 0] = w_dc2[0]; red[warp * 8 + 1] = w_dc2[1]; red[warp * 8 + 2] = w_dc2[2];
        red[warp * 8 + 3] = w_loss; red[warp * 8 + 4] = w_rec;
    }
    __syncthreads();
    if (tid < 5) {
        float t = 0.f;
        for (int w = 0; w < TR_THREADS / 32; w++) t += red[w * 8 + tid];
        const int base = Hd * ME + Hd + MS * Hd;
        if (tid < 3) P[base + tid] = t;              // dc2
        else P[base + 4 + (tid - 3)] = t;            // loss, loss_reconstruct
    }
}

// ---- 3t. the same on tensor cores (Hd <= 512) -----------------------------------------------------------------------------
// The three contractions of the decoder step over a 64-row tile — hidden = relu(g V1^T + c1) [64 x 32 x Hd],
// dV1 += dh^T g [Hd x 64 x 32] and dg = dh V1 [64 x Hd x 32] — are 3xTF32 mma.sync tile products (tct::tc_gemm, operands
// through strides: the transposed dh needs no staging); the rank-3 pieces (y = h V2^T, dh = relu'(.) (dy V2), dV2, dc1)
// stay on the FP32 pipe.  The hidden tile [64][516], the dV1 accumulator [512][36] and the g / dg tiles live in shared
// memory (220 KB, one persistent CTA per SM); hidden units >= Hd of the 512-wide tile are forced to zero, so what the
// fragment loads read beyond V1's 500 rows (the next blob section, finite weights) never contributes.
// Same per-CTA partial layout as dec_fwd_bwd_kernel, so finalize_kernel is shared.  OPT-IN: see the launch site for the
// accuracy measurement that keeps the FP32 kernel the default.
constexpr int DT_ROWS = 64, DT_HP = 512, DT_LDH = 516, DT_LDG = 36;
constexpr int DT_SMEM_FLOATS = 2 * DT_ROWS * DT_LDG + DT_ROWS * DT_LDH + DT_ROWS * 4 + DT_HP * DT_LDG + 64;

__global__ void __launch_bounds__(256, 1) dec_fwd_bwd_tc_kernel(const float* __restrict__ blob, MlpLayout L,
                                                                const float* __restrict__ states,
                                                                const float* __restrict__ G, float* __restrict__ dG,
                                                                float* __restrict__ part, TrainDims D) {
    using namespace trpx::tct;
    extern __shared__ __align__(16) float sm[];
    const int Hd = D.Hd;
    float* g_s = sm;                               // [64][36]  observables of the tile
    float* dg_s = g_s + DT_ROWS * DT_LDG;          // [64][36]  d(loss)/dg of the tile
    float* h_s = dg_s + DT_ROWS * DT_LDG;          // [64][516] hidden activations, then d(loss)/d(pre-activation)
    float* dy_s = h_s + DT_ROWS * DT_LDH;          // [64][4]
    float* dv1_s = dy_s + DT_ROWS * 4;             // [512][36] accumulated over the CTA's tiles
    float* red = dv1_s + DT_HP * DT_LDG;           // [8 warps][8]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* V1 = blob + L.decA;               // [Hd][32]
    const float* dB = blob + L.decB;               // [Hd][4] = {V2[0][u], V2[1][u], V2[2][u], c1[u]}
    for (int i = tid; i < DT_HP * DT_LDG; i += 256) dv1_s[i] = 0.f;
    float v2r[2][MS], dv2[2][MS], dc1[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int u = tid + j * 256;
#pragma unroll
        for (int c = 0; c < MS; c++) { v2r[j][c] = u < Hd ? dB[u * 4 + c] : 0.f; dv2[j][c] = 0.f; }
        dc1[j] = 0.f;
    }
    float w_dc2[MS] = {0.f, 0.f, 0.f}, w_loss = 0.f, w_rec = 0.f;     // lane-0 accumulators of each warp
    float mu[MS], sd[MS], c2[MS];
#pragma unroll
    for (int c = 0; c < MS; c++) { mu[c] = blob[L.mu + c]; sd[c] = blob[L.sd + c]; c2[c] = blob[L.c2 + c]; }

    const int ntiles = (D.M + DT_ROWS - 1) / DT_ROWS;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int r0 = tile * DT_ROWS;
        for (int i = tid; i < DT_ROWS * ME; i += 256) {
            const int s = i >> 5, k = i & 31;
            g_s[s * DT_LDG + k] = (r0 + s < D.M) ? G[(size_t)(r0 + s) * ME + k] : 0.f;
            dg_s[s * DT_LDG + k] = 0.f;
        }
        __syncthreads();
        // hidden = relu(g V1^T + c1): M = 64 rows, N = 512 units, K = 32; B(k, n = u) = V1[u][k]
        tc_gemm<2, 4, 4, true, true>(g_s, 0, DT_LDG, 1, V1, 0, 1, ME, 1, 4, 64, [&](int, int m, int n, float v) {
            h_s[m * DT_LDH + n] = n < Hd ? fmaxf(v + __ldg(dB + n * 4 + 3), 0.f) : 0.f;
        });
        __syncthreads();
        // outputs, loss terms and d(loss)/dy: 8 rows per warp, lanes over the hidden units
        for (int qq = 0; qq < DT_ROWS / 8; qq++) {
            const int s = warp * (DT_ROWS / 8) + qq;
            const int r = r0 + s;
            float y[MS] = {0.f, 0.f, 0.f};
            for (int u = lane; u < Hd; u += 32) {
                const float hv = h_s[s * DT_LDH + u];
                const float4 w = *reinterpret_cast<const float4*>(dB + u * 4);
                y[0] = fmaf(hv, w.x, y[0]);
                y[1] = fmaf(hv, w.y, y[1]);
                y[2] = fmaf(hv, w.z, y[2]);
            }
#pragma unroll
            for (int c = 0; c < MS; c++) y[c] = warp_sum(y[c]);
            if (lane == 0) {
                float dyv[MS] = {0.f, 0.f, 0.f};
                if (r < D.M) {
                    int b, t;
                    float a;
                    if (r < D.N) { b = r / D.T; t = r - b * D.T; a = D.w_recon; }
                    else { const int rr = r - D.N; b = rr / (D.T - 1); t = rr - b * (D.T - 1) + 1; a = D.w_dyn; }
                    const float* xt = states + ((size_t)b * D.T + t) * MS;
                    float sq = 0.f;
#pragma unroll
                    for (int c = 0; c < MS; c++) {
                        const float xh = sd[c] * (y[c] + c2[c]) + mu[c];
                        const float df = xh - xt[c];
                        sq = fmaf(df, df, sq);
                        dyv[c] = sd[c] * (a * 2.0f * df * D.inv3B);
                        w_dc2[c] += dyv[c];
                    }
                    w_loss += a * sq * D.inv3B;
                    if (r < D.N) w_rec += sq * D.inv3B;
                }
                *reinterpret_cast<float4*>(dy_s + s * 4) = make_float4(dyv[0], dyv[1], dyv[2], 0.f);
            }
        }
        __syncthreads();
        // per-unit rank-3 backward: dV2, dc1, and h_s := d(loss)/d(hidden pre-activation)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int u = tid + j * 256;
            if (u < Hd) {
#pragma unroll 4
                for (int s = 0; s < DT_ROWS; s++) {
                    const float hv = h_s[s * DT_LDH + u];
                    const float4 dyv = *reinterpret_cast<const float4*>(dy_s + s * 4);
                    dv2[j][0] = fmaf(dyv.x, hv, dv2[j][0]);
                    dv2[j][1] = fmaf(dyv.y, hv, dv2[j][1]);
                    dv2[j][2] = fmaf(dyv.z, hv, dv2[j][2]);
                    float dh = fmaf(v2r[j][2], dyv.z, fmaf(v2r[j][1], dyv.y, v2r[j][0] * dyv.x));
                    dh = hv > 0.f ? dh : 0.f;
                    dc1[j] += dh;
                    h_s[s * DT_LDH + u] = dh;
                }
            }
        }
        __syncthreads();
        // dV1 += dh^T g  (M = 512 units, N = 32, K = 64 rows; A transposed through its strides)
        tc_gemm<2, 4, 8, false, false>(h_s, 0, 1, DT_LDH, g_s, 0, DT_LDG, 1, 1, 32, 4,
                                       [&](int, int m, int n, float v) { dv1_s[m * DT_LDG + n] += v; });
        // dg = dh V1  (M = 64, N = 32, K = 512 in 8 chunks of 64; B(k = u, n) = V1[u][n])
        for (int kc = 0; kc < DT_HP / 64; kc++)
            tc_gemm<1, 2, 8, true, false>(h_s + kc * 64, 0, DT_LDH, 1, V1 + (size_t)kc * 64 * ME, 0, ME, 1, 1, 4, 4,
                                          [&](int, int m, int n, float v) { dg_s[m * DT_LDG + n] += v; });
        __syncthreads();
        for (int i = tid; i < DT_ROWS * 8; i += 256) {
            const int s = i >> 3, k4 = (i & 7) * 4;
            if (r0 + s < D.M)
                *reinterpret_cast<float4*>(dG + (size_t)(r0 + s) * ME + k4) = *reinterpret_cast<const float4*>(dg_s + s * DT_LDG + k4);
        }
        __syncthreads();
    }
    // per-CTA partials
    float* P = part + (size_t)blockIdx.x * dec_part_size(Hd);
    for (int i = tid; i < Hd * ME; i += 256) P[i] = dv1_s[(i >> 5) * DT_LDG + (i & 31)];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int u = tid + j * 256;
        if (u < Hd) {
            P[Hd * ME + u] = dc1[j];
#pragma unroll
            for (int c = 0; c < MS; c++) P[Hd * ME + Hd + c * Hd + u] = dv2[j][c];
        }
    }
    if (lane == 0) {
        red[warp * 8 + 0] = w_dc2[0]; red[warp * 8 + 1] = w_dc2[1]; red[warp * 8 + 2] = w_dc2[2];
        red[warp * 8 + 3] = w_loss; red[warp * 8 + 4] = w_rec;
    }
    __syncthreads();
    if (tid < 5) {
        float t = 0.f;
        for (int w = 0; w < 8; w++) t += red[w * 8 + tid];
        const int base = Hd * ME + Hd + MS * Hd;
        if (tid < 3) P[base + tid] = t;              // dc2
        else P[base + 4 + (tid - 3)] = t;            // loss, loss_reconstruct
    }
}

// ---- 4. back through the Koopman recurrence ------------------------------------------------------------
// one warp per trajectory; lane i accumulates d/d kMatrixDiag[i], d/d u1[i], d/d u2[i]; the carry reaching
// t = 0 is added to d(loss)/dg of the encoder sample (b, 0).
__global__ void chain_bwd_kernel(const float* __restrict__ blob, MlpLayout L, const float* __restrict__ G,
                                 float* __restrict__ dG, float* __restrict__ part, TrainDims D) {
    __shared__ float red[8][3][ME];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const BandRegs k = load_band(blob, L, lane);
    float ad = 0.f, a1 = 0.f, a2 = 0.f;
    for (int b = warp; b < D.B; b += nwarps) {
        float carry = 0.f;
        for (int t = D.T - 1; t >= 1; t--) {
            const size_t row = (size_t)D.N + (size_t)b * (D.T - 1) + (t - 1);
            const float dg = dG[row * ME + lane] + carry;
            const size_t prow = (t == 1) ? (size_t)b * D.T : row - 1;
            const float gp = G[prow * ME + lane];
            const float dg_p1 = __shfl_down_sync(FULL, dg, 1), dg_p2 = __shfl_down_sync(FULL, dg, 2);
            const float dg_m1 = __shfl_up_sync(FULL, dg, 1), dg_m2 = __shfl_up_sync(FULL, dg, 2);
            const float gp_p1 = __shfl_down_sync(FULL, gp, 1), gp_p2 = __shfl_down_sync(FULL, gp, 2);
            ad = fmaf(dg, gp, ad);
            if (lane < ME - 1) a1 += dg * gp_p1 - dg_p1 * gp;      // K[i][i+1] = u1[i], K[i+1][i] = -u1[i]
            if (lane < ME - 2) a2 += dg * gp_p2 - dg_p2 * gp;
            // carry[j] = sum_i K[i][j] dg[i]
            float c = k.d * dg;
            if (lane >= 1) c = fmaf(k.u1m, dg_m1, c);
            if (lane < ME - 1) c = fmaf(-k.u1, dg_p1, c);
            if (lane >= 2) c = fmaf(k.u2m, dg_m2, c);
            if (lane < ME - 2) c = fmaf(-k.u2, dg_p2, c);
            carry = c;
        }
        if (D.T > 1) dG[((size_t)b * D.T) * ME + lane] += carry;
    }
    red[wib][0][lane] = ad; red[wib][1][lane] = a1; red[wib][2][lane] = a2;
    __syncthreads();
    if (threadIdx.x < 3 * ME) {
        const int which = threadIdx.x / ME, i = threadIdx.x % ME;
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w][which][i];
        part[(size_t)blockIdx.x * 3 * ME + which * ME + i] = t;
    }
}

// ---- 5. LayerNorm + encoder backward ---------------------------------------------------------------------
__global__ void __launch_bounds__(TR_THREADS, 1) enc_bwd_kernel(const float* __restrict__ blob, MlpLayout L,
                                                                const float* __restrict__ states,
                                                                const float* __restrict__ zhat,
                                                                const float* __restrict__ rstd,
                                                                const float* __restrict__ dG, float* __restrict__ part,
                                                                TrainDims D) {
    __shared__ __align__(16) float dz_s[TILE][ME];
    __shared__ __align__(16) float xn_s[TILE][4];
    __shared__ float red[8][3][ME];
    const int Hd = D.Hd;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* eA = blob + L.encA;      // [Hd][4] = {W1[u][0..2], b1[u]}
    const float* eB = blob + L.encB;      // [Hd][32] = W2[e][u]
    float w2c[NU][ME], dw2[NU][ME], w1r[NU][4], dw1[NU][MS], db1[NU];
#pragma unroll
    for (int j = 0; j < NU; j++) {
        const int u = tid + j * TR_THREADS;
        const bool ok = u < Hd;
#pragma unroll
        for (int e = 0; e < ME; e++) { w2c[j][e] = ok ? eB[u * ME + e] : 0.f; dw2[j][e] = 0.f; }
#pragma unroll
        for (int q = 0; q < 4; q++) w1r[j][q] = ok ? eA[u * 4 + q] : 0.f;
#pragma unroll
        for (int c = 0; c < MS; c++) dw1[j][c] = 0.f;
        db1[j] = 0.f;
    }
    const float lnw = blob[L.lnw + lane];
    float a_lnw = 0.f, a_lnb = 0.f, a_b2 = 0.f;       // lane = embedding channel
    const int ntiles = (D.N + TILE - 1) / TILE;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int r0 = tile * TILE;
        // LayerNorm backward, 4 samples per warp
        for (int q = 0; q < TILE / 8; q++) {
            const int s = warp * (TILE / 8) + q;
            const int r = r0 + s;
            float dz = 0.f;
            if (r < D.N) {
                const float dgv = dG[(size_t)r * ME + lane];
                const float zh = zhat[(size_t)r * ME + lane];
                const float dgh = dgv * lnw;
                const float s1 = warp_sum(dgh), s2 = warp_sum(dgh * zh);
                dz = rstd[r] * (dgh - s1 * (1.0f / ME) - zh * (s2 * (1.0f / ME)));
                a_lnw = fmaf(dgv, zh, a_lnw);
                a_lnb += dgv;
                a_b2 += dz;
            }
            dz_s[s][lane] = dz;
            if (lane < 4) {
                float v = 0.f;
                if (r < D.N && lane < MS) v = (states[(size_t)r * MS + lane] - blob[L.mu + lane]) / blob[L.sd + lane];
                xn_s[s][lane] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NU; j++) {
            const int u = tid + j * TR_THREADS;
            if (u < Hd) {
                for (int s = 0; s < TILE; s++) {
                    const float4 xn = *reinterpret_cast<const float4*>(&xn_s[s][0]);
                    const float pre = fmaf(w1r[j][2], xn.z, fmaf(w1r[j][1], xn.y, fmaf(w1r[j][0], xn.x, w1r[j][3])));
                    const float h1 = fmaxf(pre, 0.f);
                    const float4* dzr = reinterpret_cast<const float4*>(&dz_s[s][0]);
                    float d0 = 0.f, d1 = 0.f;
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 dv = dzr[q];
                        d0 = fmaf(w2c[j][4 * q + 0], dv.x, d0);
                        d1 = fmaf(w2c[j][4 * q + 1], dv.y, d1);
                        d0 = fmaf(w2c[j][4 * q + 2], dv.z, d0);
                        d1 = fmaf(w2c[j][4 * q + 3], dv.w, d1);
                        dw2[j][4 * q + 0] = fmaf(dv.x, h1, dw2[j][4 * q + 0]);
                        dw2[j][4 * q + 1] = fmaf(dv.y, h1, dw2[j][4 * q + 1]);
                        dw2[j][4 * q + 2] = fmaf(dv.z, h1, dw2[j][4 * q + 2]);
                        dw2[j][4 * q + 3] = fmaf(dv.w, h1, dw2[j][4 * q + 3]);
                    }
                    const float dh = pre > 0.f ? d0 + d1 : 0.f;
                    db1[j] += dh;
                    dw1[j][0] = fmaf(dh, xn.x, dw1[j][0]);
                    dw1[j][1] = fmaf(dh, xn.y, dw1[j][1]);
                    dw1[j][2] = fmaf(dh, xn.z, dw1[j][2]);
                }
            }
        }
        __syncthreads();
    }
    float* P = part + (size_t)blockIdx.x * enc_part_size(Hd);
#pragma unroll
    for (int j = 0; j < NU; j++) {
        const int u = tid + j * TR_THREADS;
        if (u < Hd) {
#pragma unroll
            for (int c = 0; c < MS; c++) P[u * MS + c] = dw1[j][c];
            P[Hd * MS + u] = db1[j];
#pragma unroll
            for (int e = 0; e < ME; e++) P[Hd * MS + Hd + e * Hd + u] = dw2[j][e];
        }
    }
    red[warp][0][lane] = a_b2; red[warp][1][lane] = a_lnw; red[warp][2][lane] = a_lnb;
    __syncthreads();
    if (tid < 3 * ME) {
        const int which = tid / ME, e = tid % ME;
        float t = 0.f;
        for (int w = 0; w < TR_THREADS / 32; w++) t += red[w][which][e];
        P[Hd * MS + Hd + ME * Hd + which * ME + e] = t;
    }
}

// ---- 6. ordered reduction of the partials into the flat gradient + losses --------------------------------
struct FinalArgs {
    const float *dec_part, *enc_part, *chain_part;
    int n_dec, n_enc, n_chain;
    float grad_scale;
};

// sum of n per-CTA partials (stride floats apart) in a FIXED order: four interleaved chains (c mod 4), then
// (s0 + s1) + (s2 + s3) — deterministic, and four loads in flight instead of one dependent load-add per partial
__device__ __forceinline__ float sum_partials(const float* __restrict__ p, size_t stride, int n) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int c = 0;
    for (; c + 4 <= n; c += 4) {
        s0 += p[(size_t)c * stride];
        s1 += p[(size_t)(c + 1) * stride];
        s2 += p[(size_t)(c + 2) * stride];
        s3 += p[(size_t)(c + 3) * stride];
    }
    for (; c < n; c++) s0 += p[(size_t)c * stride];
    return (s0 + s1) + (s2 + s3);
}

__global__ void finalize_kernel(const float* __restrict__ blob, MlpLayout L, FinalArgs A, TrainDims D,
                                float* __restrict__ grad, float* __restrict__ loss2) {
    const int Hd = D.Hd;
    const int DP = dec_part_size(Hd), EP = enc_part_size(Hd);
    const int n_total = ME + 61 + Hd * MS + Hd + ME * Hd + 3 * ME + Hd * ME + Hd + MS * Hd + MS;
    const float regT = D.w_reg * (float)(D.T - 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n_total; i += gridDim.x * blockDim.x) {
        if (i == n_total) {
            // losses (thread handling the virtual index n_total)
            float ls = 0.f, lr = 0.f;
            for (int c = 0; c < A.n_dec; c++) {
                ls += A.dec_part[(size_t)c * DP + Hd * ME + Hd + MS * Hd + 4];
                lr += A.dec_part[(size_t)c * DP + Hd * ME + Hd + MS * Hd + 5];
            }
            float k2 = 0.f;
            for (int j = 0; j < ME; j++) k2 = fmaf(blob[L.kd + j], blob[L.kd + j], k2);
            for (int j = 0; j < 61; j++) k2 = fmaf(2.0f * blob[L.ku + j], blob[L.ku + j], k2);
            loss2[0] = ls + regT * k2;
            loss2[1] = lr;
            continue;
        }
        int o = i;
        float v = 0.f;
        if (o < ME) {                                   // kMatrixDiag
            for (int c = 0; c < A.n_chain; c++) v += A.chain_part[(size_t)c * 3 * ME + o];
            v += 2.0f * regT * blob[L.kd + o];
        } else if ((o -= ME) < 61) {                    // kMatrixUT: 31 offset-1 entries then 30 offset-2 entries
            const int which = o < ME - 1 ? 1 : 2, idx = o < ME - 1 ? o : o - (ME - 1);
            for (int c = 0; c < A.n_chain; c++) v += A.chain_part[(size_t)c * 3 * ME + which * ME + idx];
            v += 4.0f * regT * blob[L.ku + o];
        } else if ((o -= 61) < Hd * MS + Hd) {          // observableNet.0.{weight [Hd,3], bias [Hd]}
            v = sum_partials(A.enc_part + o, EP, A.n_enc);
        } else if ((o -= Hd * MS + Hd) < ME * Hd + 3 * ME) {   // observableNet.2.{weight [32,Hd], bias}, .3.{weight, bias}
            v = sum_partials(A.enc_part + Hd * MS + Hd + o, EP, A.n_enc);
        } else {                                        // recoveryNet.0.{weight [Hd,32], bias}, .2.{weight [3,Hd], bias}
            o -= ME * Hd + 3 * ME;
            v = sum_partials(A.dec_part + o, DP, A.n_dec);
        }
        grad[i] = v * A.grad_scale;
    }
}

}  // namespace

extern "C" int trpx_mlpemb_train_fwd_bwd(trpx_mlpemb_t h, const float* states, int B, int T, float w_recon,
                                         float w_dyn, float w_reg, float grad_scale, float* loss2, float* grad,
                                         void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "weights not loaded");
    TRPX_REQUIRE(states && loss2 && grad, "null tensor");
    TRPX_REQUIRE(B >= 1 && T >= 1, "B and T must be >= 1");
    TRPX_REQUIRE(h->cfg.hidden <= NU * TR_THREADS, "hidden=%d too wide for the training kernels", h->cfg.hidden);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const MlpLayout L = MlpLayout::make(h->cfg.hidden);
    TrainDims D;
    D.B = B; D.T = T; D.N = B * T; D.M = D.N + B * (T - 1); D.Hd = h->cfg.hidden;
    D.w_recon = w_recon; D.w_dyn = w_dyn; D.w_reg = w_reg; D.inv3B = 1.0f / (float)(MS * B);
    const int n_dec = std::min(h->sm_count, (D.M + TILE - 1) / TILE);
    const int n_enc = std::min(h->sm_count, (D.N + TILE - 1) / TILE);
    const int n_chain = std::min(64, (B + 7) / 8);
    const int DP = dec_part_size(D.Hd), EP = enc_part_size(D.Hd);
    // workspace carve-up (floats)
    size_t o = 0;
    const size_t oG = o; o += (size_t)D.M * ME;
    const size_t oZ = o; o += (size_t)D.N * ME;
    const size_t oR = o; o += ((size_t)D.N + 3) & ~(size_t)3;
    const size_t oDG = o; o += (size_t)D.M * ME;
    const size_t oPD = o; o += (size_t)n_dec * DP;
    const size_t oPE = o; o += (size_t)n_enc * EP;
    const size_t oPC = o; o += (size_t)n_chain * 3 * ME;
    if (o * sizeof(float) > h->ws_bytes) {
        if (h->ws) { TRPX_CUDA(cudaDeviceSynchronize()); TRPX_CUDA(cudaFree(h->ws)); h->ws = nullptr; h->ws_bytes = 0; }
        TRPX_CUDA(cudaMalloc(&h->ws, o * sizeof(float)));
        h->ws_bytes = o * sizeof(float);
    }
    float* W = h->ws;
    int rc = mlpemb_embed_ex(h, states, W + oG, D.N, W + oZ, W + oR, stream);
    if (rc) return rc;
    if (T > 1) {
        chain_fwd_kernel<<<n_chain, 256, 0, st>>>(h->blob, L, W + oG, D);
        TRPX_CUDA(cudaGetLastError());
    }
    {
        // The tensor-core kernel is OPT-IN (TRPX_KOOPMAN_TC=1): it takes the step from 0.25 to 0.19 ms, but this gradient is
        // ill conditioned (loss weights 1e4, sums over 15,872 rows that cancel ~50x): against an FP64 evaluation the FP32
        // kernel is at 4.1e-6 and the 3xTF32 products (2^-22 per product, intrinsic) at 1.06e-5 — over the 1e-5 parity bar.
        const char* tc = getenv("TRPX_KOOPMAN_TC");
        const size_t smem_tc = sizeof(float) * (size_t)DT_SMEM_FLOATS;
        if (D.Hd <= DT_HP && tc && tc[0] == '1' && smem_tc + 1024 <= (size_t)h->smem_optin) {
            TRPX_CUDA(cudaFuncSetAttribute(dec_fwd_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
            dec_fwd_bwd_tc_kernel<<<n_dec, 256, smem_tc, st>>>(h->blob, L, states, W + oG, W + oDG, W + oPD, D);
        } else {
            const size_t smem = sizeof(float) * ((size_t)TILE * ME + (size_t)TILE * D.Hd + TILE * 4 + 64);
            TRPX_CUDA(cudaFuncSetAttribute(dec_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dec_fwd_bwd_kernel<<<n_dec, TR_THREADS, smem, st>>>(h->blob, L, states, W + oG, W + oDG, W + oPD, D);
        }
    }
    TRPX_CUDA(cudaGetLastError());
    chain_bwd_kernel<<<n_chain, 256, 0, st>>>(h->blob, L, W + oG, W + oDG, W + oPC, D);
    TRPX_CUDA(cudaGetLastError());
    enc_bwd_kernel<<<n_enc, TR_THREADS, 0, st>>>(h->blob, L, states, W + oZ, W + oR, W + oDG, W + oPE, D);
    TRPX_CUDA(cudaGetLastError());
    FinalArgs A{W + oPD, W + oPE, W + oPC, n_dec, n_enc, n_chain, grad_scale};
    finalize_kernel<<<72, 512, 0, st>>>(h->blob, L, A, D, grad, loss2);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}
