// K8: teacher-forced training step of the decoder stack.
//
// Replaces, for one mini-batch, PhysformerTrain.forward + loss.backward()
// (trphysx/transformer/phys_transformer_helpers.py:36-69 ; utils/trainer.py:244-277):
//     hidden = transformer.forward(inputs_embeds)[0];  loss = MSELoss()(hidden, labels_embeds)   (mean over B*T*E)
// and the gradient of `loss` with respect to every parameter of the stack, in the packed blob layout
// (gpt2_layout.cuh; mlp_f.weight transposed).  Forward arithmetic is the one of the forward kernel (K2).
//
// One CTA per sequence: forward keeps the residual stream in shared memory and saves, per layer, the tensors the
// backward needs (LN inputs/outputs, q|k|v, attention output, MLP pre-activation) in a global workspace that stays
// L2 resident (12*T*E floats per layer and sequence); attention probabilities are recomputed per head in the
// backward.  Weight gradients are written as per-sequence partials and summed over the batch in a fixed order by
// the finalize kernel (deterministic; no atomics), which also finishes the loss.
#include "common.cuh"
#include "gpt2_layout.cuh"
#include "gpt2_handle.cuh"
#include "gpt2_train_tc.cuh"
#include <algorithm>
#include <cstdlib>

using namespace trpx;

namespace {

constexpr int TR_THREADS = 512;      // 128 registers per thread for the register-blocked tile products

struct TrainParams {
    const float* blob;
    const float* pe;
    const float* x;         // [B][T][E]
    const float* labels;    // [B][T][E]
    float* hidden;          // [B][T][E] or null
    float* acts;            // [B][acts_per_seq]
    float* gpart;           // [B][L.total]
    float* loss_part;       // [B]
    int B, T;
    float eps;
    float dscale;           // 2 / (B*T*E)
    Gpt2Layout lay;
};

__device__ __forceinline__ float gelu_new_grad(float x) {
    // d/dx [0.5 x (1 + tanh(u))], u = c (x + 0.044715 x^3)     (transformer/utils.py:57-60)
    const float c = 0.7978845608028654f;
    const float u = c * (x + 0.044715f * x * x * x);
    const float t = tanhf(u);
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * c * (1.0f + 3.0f * 0.044715f * x * x);
}

// The three tile products below are register-blocked over 4 tokens (or 4 rows of the contraction) and read the smem
// operand with 128-bit loads: ~1.5 instructions per FMA instead of ~5 (the kernel was issue bound: 25 % of the warp
// samples "not selected" with 32 warps per SM).  Every output is still accumulated in ascending contraction order
// starting from its bias, so results are bit-identical to the plain loops.

// Y[t][n] = bias[n] + sum_k X[t][k] * W[k*N + n]            X: smem [T][K] (K % 4 == 0), W: global [K][N]
__device__ void mm_xw(float* Y, const float* X, const float* __restrict__ W, const float* __restrict__ bias, int T,
                      int K, int N, int ldy = 0) {
    if (ldy == 0) ldy = N;
    const int TB = (T + 3) >> 2;
    for (int u = threadIdx.x; u < TB * N; u += blockDim.x) {
        const int tb = u / N, n = u - tb * N;
        const float b = bias ? __ldg(bias + n) : 0.f;
        float acc[4] = {b, b, b, b};
        const float* xr[4];
#pragma unroll
        for (int i = 0; i < 4; i++) xr[i] = X + min(4 * tb + i, T - 1) * K;
#pragma unroll 2
        for (int k = 0; k < K; k += 4) {
            const float* wp = W + (size_t)k * N + n;
            const float w0 = __ldg(wp), w1 = __ldg(wp + N), w2 = __ldg(wp + 2 * N), w3 = __ldg(wp + 3 * N);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 x = *reinterpret_cast<const float4*>(xr[i] + k);
                acc[i] = fmaf(x.x, w0, acc[i]);
                acc[i] = fmaf(x.y, w1, acc[i]);
                acc[i] = fmaf(x.z, w2, acc[i]);
                acc[i] = fmaf(x.w, w3, acc[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (4 * tb + i < T) Y[(4 * tb + i) * ldy + n] = acc[i];
    }
}
// dX[t][k] = sum_n dY[t][n] * W[k*N + n]                     dY: smem [T][N] (N % 4 == 0)
// W rows are contiguous along n, the outputs want lanes along k: column chunks of W are staged TRANSPOSED in shared
// memory (`stage`, STAGE_FLOATS floats, row stride K+1: conflict-free for both the transposing store and the read),
// so every global read is coalesced.  A thread owns up to MAXU units of (4 tokens, k) and keeps them in registers
// across the chunks.  All threads must call; synchronises internally.
constexpr int STAGE_FLOATS = 8192 + 512;
template <int MAXU>
__device__ void mm_xwT(float* dX, const float* dY, const float* __restrict__ W, int T, int K, int N, float* stage) {
    const int NC = min(N, 8192 / K);                 // columns per chunk (a multiple of 4)
    const int TB = (T + 3) >> 2;
    float acc[MAXU][4];
#pragma unroll
    for (int o = 0; o < MAXU; o++) acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f;
    for (int n0 = 0; n0 < N; n0 += NC) {
        const int nc = min(NC, N - n0);
        __syncthreads();                              // previous chunk fully consumed
        for (int i = threadIdx.x; i < K * nc; i += blockDim.x) {
            const int k = i / nc, nn = i - k * nc;
            stage[nn * (K + 1) + k] = __ldg(W + (size_t)k * N + n0 + nn);
        }
        __syncthreads();
#pragma unroll
        for (int o = 0; o < MAXU; o++) {
            const int u = threadIdx.x + o * blockDim.x;
            if (u < TB * K) {
                const int tb = u / K, k = u - tb * K;
                const float* dr[4];
#pragma unroll
                for (int i = 0; i < 4; i++) dr[i] = dY + min(4 * tb + i, T - 1) * N + n0;
#pragma unroll 2
                for (int nn = 0; nn < nc; nn += 4) {
                    const float* sp = stage + nn * (K + 1) + k;
                    const float w0 = sp[0], w1 = sp[K + 1], w2 = sp[2 * (K + 1)], w3 = sp[3 * (K + 1)];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float4 d = *reinterpret_cast<const float4*>(dr[i] + nn);
                        acc[o][i] = fmaf(d.x, w0, acc[o][i]);
                        acc[o][i] = fmaf(d.y, w1, acc[o][i]);
                        acc[o][i] = fmaf(d.z, w2, acc[o][i]);
                        acc[o][i] = fmaf(d.w, w3, acc[o][i]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < MAXU; o++) {
        const int u = threadIdx.x + o * blockDim.x;
        if (u < TB * K) {
            const int tb = u / K, k = u - tb * K;
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (4 * tb + i < T) dX[(4 * tb + i) * K + k] = acc[o][i];
        }
    }
}
// dW[k*N + n] = sum_t X[t][k] * dY[t][n] ; db[n] = sum_t dY[t][n]      (written, not accumulated; K % 4 == 0)
__device__ void mm_xTy(float* dW, float* db, const float* X, const float* dY, int T, int K, int N) {
    for (int u = threadIdx.x; u < (K >> 2) * N; u += blockDim.x) {
        const int kb = u / N, n = u - kb * N;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
        for (int t = 0; t < T; t++) {
            const float dy = dY[t * N + n];
            const float4 x = *reinterpret_cast<const float4*>(X + t * K + 4 * kb);
            a0 = fmaf(x.x, dy, a0);
            a1 = fmaf(x.y, dy, a1);
            a2 = fmaf(x.z, dy, a2);
            a3 = fmaf(x.w, dy, a3);
        }
        float* o = dW + (size_t)(4 * kb) * N + n;
        o[0] = a0; o[N] = a1; o[2 * N] = a2; o[3 * N] = a3;
    }
    if (db != nullptr)
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            float acc = 0.f;
            for (int t = 0; t < T; t++) acc += dY[t * N + n];
            db[n] = acc;
        }
}
// Y = LN(X) (biased variance)
__device__ void ln_fwd(float* Y, const float* X, const float* __restrict__ w, const float* __restrict__ b, int T, int E,
                       float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int t = warp; t < T; t += nw) {
        float s = 0.f;
        for (int c = lane; c < E; c += 32) s += X[t * E + c];
        const float mean = warp_sum(s) / (float)E;
        float v = 0.f;
        for (int c = lane; c < E; c += 32) {
            const float d = X[t * E + c] - mean;
            v = fmaf(d, d, v);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)E + eps);
        for (int c = lane; c < E; c += 32) Y[t * E + c] = (X[t * E + c] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
    }
}
// LayerNorm backward: X (in: LN input, out: xhat), dY unchanged; dH[t][c] (+)= dx ; dw, db written.
// Caller synchronises before (X, dY complete) — this function synchronises internally between its two phases.
__device__ void ln_bwd(float* dH, bool accumulate, float* X, const float* dY, const float* __restrict__ w, float* dw,
                       float* db, int T, int E, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int t = warp; t < T; t += nw) {
        float s = 0.f;
        for (int c = lane; c < E; c += 32) s += X[t * E + c];
        const float mean = warp_sum(s) / (float)E;
        float v = 0.f;
        for (int c = lane; c < E; c += 32) {
            const float d = X[t * E + c] - mean;
            v = fmaf(d, d, v);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)E + eps);
        float m1 = 0.f, m2 = 0.f;
        for (int c = lane; c < E; c += 32) {
            const float xh = (X[t * E + c] - mean) * rstd;
            const float dyw = dY[t * E + c] * __ldg(w + c);
            X[t * E + c] = xh;
            m1 += dyw;
            m2 = fmaf(dyw, xh, m2);
        }
        m1 = warp_sum(m1) / (float)E;
        m2 = warp_sum(m2) / (float)E;
        for (int c = lane; c < E; c += 32) {
            const float dyw = dY[t * E + c] * __ldg(w + c);
            const float dx = rstd * (dyw - m1 - X[t * E + c] * m2);
            dH[t * E + c] = accumulate ? dH[t * E + c] + dx : dx;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < E; c += blockDim.x) {
        float a = 0.f, bsum = 0.f;
        for (int t = 0; t < T; t++) {
            a = fmaf(dY[t * E + c], X[t * E + c], a);
            bsum += dY[t * E + c];
        }
        dw[c] = a;
        db[c] = bsum;
    }
}
__device__ void tile_load(float* dst, const float* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
__device__ void tile_store(float* __restrict__ dst, const float* src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// [T][n] <-> [T][ld] (padded rows)
__device__ void tile_load_ld(float* dst, int ld, const float* __restrict__ src, int T, int n) {
    for (int i = threadIdx.x; i < T * n; i += blockDim.x) dst[(i / n) * ld + (i % n)] = src[i];
}
__device__ void tile_store_ld(float* __restrict__ dst, const float* src, int ld, int T, int n) {
    for (int i = threadIdx.x; i < T * n; i += blockDim.x) dst[i] = src[(i / n) * ld + (i % n)];
}

// causal softmax row i of head hh (q|k|v tile with row stride LQ = 3E+1: lanes walk rows j, so an unpadded stride of
// 3E = 0 mod 32 would put every lane on the same bank):  P[i][j] = softmax_{j<=i}(q_i . k_j / sqrt(hd)), 0 for j > i
// (attention.py:97-109: masked logits are -1e4, whose exp underflows to exactly 0 in FP32)
__device__ __forceinline__ void attn_row(const float* qkv, int i, int hh, int T, int E, int hd, float* Prow, int lane) {
    const float scale = sqrtf((float)hd);
    const int LQ = 3 * E + 1;
    const float* q = qkv + i * LQ + hh * hd;
    float m = -INFINITY;
    for (int j = lane; j <= i; j += 32) {
        const float* k = qkv + j * LQ + E + hh * hd;
        float s = 0.f;
        for (int d = 0; d < hd; d++) s = fmaf(q[d], k[d], s);
        s = s / scale;
        Prow[j] = s;
        m = fmaxf(m, s);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j <= i; j += 32) {
        const float e = expf(Prow[j] - m);
        Prow[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    for (int j = lane; j < T; j += 32) Prow[j] = (j <= i) ? Prow[j] / sum : 0.f;
    __syncwarp();
}

__global__ void __launch_bounds__(TR_THREADS) train_fwd_bwd_kernel(TrainParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float red[TR_THREADS / 32];
    const Gpt2Layout& L = p.lay;
    const int E = L.E, H = L.n_head, hd = L.hd, T = p.T, NL = L.n_layer;
    const int TE = T * E;
    const int LQ = 3 * E + 1;              // row stride of the q|k|v tile (see attn_row)
    float* h = sm;                         // [T][E]   residual (forward) / its gradient (backward)
    float* t0 = h + TE;                    // four [T][4E] work tiles
    float* t1 = t0 + 4 * TE;
    float* t2 = t1 + 4 * TE;
    float* t3 = t2 + 4 * TE;
    float* P = t3 + 4 * TE;                // [max(T, warps)][T]  (forward: one scratch row per warp)
    float* dS = P + max(T, TR_THREADS / 32) * T;   // [T][T]
    float* stage = dS + T * T;             // [STAGE_FLOATS] transposed weight chunks (mm_xwT)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    const int b = blockIdx.x;
    const size_t per_layer = (size_t)12 * TE;
    float* acts = p.acts + (size_t)b * (NL * per_layer + 2 * (size_t)TE);
    float* gp = p.gpart + (size_t)b * L.total;
    // saved per layer: x_in [TE] | a1 [TE] | qkv [3TE] | att [TE] | x_mid [TE] | a2 [TE] | fc_pre [4TE]
    constexpr int O_XIN = 0, O_A1 = 1, O_QKV = 2, O_ATT = 5, O_XMID = 6, O_A2 = 7, O_FC = 8;

    // ============================== forward ==============================
    for (int i = tid; i < TE; i += blockDim.x) {
        const int t = i / E, c = i - t * E;
        h[i] = p.x[(size_t)b * TE + i] + __ldg(p.pe + t * E + c);      // position id = t (phys_transformer_gpt2.py:198)
    }
    __syncthreads();
    for (int l = 0; l < NL; l++) {
        const float* w = p.blob + (size_t)l * L.layer_stride;
        float* sv = acts + l * per_layer;
        tile_store(sv + O_XIN * TE, h, TE);
        ln_fwd(t0, h, w + L.ln1w, w + L.ln1b, T, E, p.eps);
        __syncthreads();
        tile_store(sv + O_A1 * TE, t0, TE);
        mm_xw(t1, t0, w + L.wqkv, w + L.bqkv, T, E, 3 * E, LQ);        // [T][LQ] rows q|k|v (padded)
        __syncthreads();
        tile_store_ld(sv + O_QKV * TE, t1, LQ, T, 3 * E);
        for (int pair = warp; pair < T * H; pair += nw) {
            const int i = pair / H, hh = pair - i * H;
            float* Prow = P + (size_t)warp * T;                         // per-warp scratch row (nw <= T rows used)
            attn_row(t1, i, hh, T, E, hd, Prow, lane);
            // o_i = P_i . V: lanes = (key part, channel) so that all 32 lanes work for small heads; parts combined by shuffles
            const int ld = hd < 32 ? hd : 32, nsp = 32 / ld, sp = lane / ld;
            for (int d = lane % ld; d < hd; d += ld) {
                float acc = 0.f;
                for (int j = sp; j <= i; j += nsp) acc = fmaf(Prow[j], t1[j * LQ + 2 * E + hh * hd + d], acc);
                for (int o = ld; o < 32; o <<= 1) acc += __shfl_xor_sync(FULL, acc, o);
                if (sp == 0) t2[i * E + hh * hd + d] = acc;
            }
            __syncwarp();
        }
        __syncthreads();
        tile_store(sv + O_ATT * TE, t2, TE);
        mm_xw(t0, t2, w + L.wproj, w + L.bproj, T, E, E);
        __syncthreads();
        for (int i = tid; i < TE; i += blockDim.x) h[i] += t0[i];
        __syncthreads();
        tile_store(sv + O_XMID * TE, h, TE);
        ln_fwd(t0, h, w + L.ln2w, w + L.ln2b, T, E, p.eps);
        __syncthreads();
        tile_store(sv + O_A2 * TE, t0, TE);
        mm_xw(t1, t0, w + L.wfc, w + L.bfc, T, E, 4 * E);
        __syncthreads();
        tile_store(sv + O_FC * TE, t1, 4 * TE);
        for (int i = tid; i < 4 * TE; i += blockDim.x) t2[i] = gelu_new(t1[i]);
        __syncthreads();
        mm_xw(t0, t2, w + L.wpr, w + L.bpr, T, 4 * E, E);
        __syncthreads();
        for (int i = tid; i < TE; i += blockDim.x) h[i] += t0[i];
        __syncthreads();
    }
    float* sv_f = acts + NL * per_layer;                                // h_L [TE] | a_f [TE]
    tile_store(sv_f, h, TE);
    ln_fwd(t0, h, p.blob + L.lnfw, p.blob + L.lnfb, T, E, p.eps);
    __syncthreads();
    mm_xw(t1, t0, p.blob + L.wf, p.blob + L.bf, T, E, E);              // blob wf is [in][out]
    __syncthreads();
    float lsum = 0.f;
    for (int i = tid; i < TE; i += blockDim.x) {
        const float y = t1[i];
        if (p.hidden) p.hidden[(size_t)b * TE + i] = y;
        const float d = y - p.labels[(size_t)b * TE + i];
        lsum = fmaf(d, d, lsum);
        t1[i] = d * p.dscale;                                           // dL/dy
    }
    lsum = warp_sum(lsum);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int i = 0; i < nw; i++) s += red[i];
        p.loss_part[b] = s;
    }

    // ============================== backward ==============================
    // head: y = a_f Wf + bf ; a_f = LN_f(h_L)
    mm_xTy(gp + L.wf, gp + L.bf, t0, t1, T, E, E);
    mm_xwT<1>(t2, t1, p.blob + L.wf, T, E, E, stage);                   // d a_f
    tile_load(t3, sv_f, TE);                                            // h_L
    __syncthreads();
    ln_bwd(h, false, t3, t2, p.blob + L.lnfw, gp + L.lnfw, gp + L.lnfb, T, E, p.eps);
    __syncthreads();
    for (int l = NL - 1; l >= 0; l--) {
        const float* w = p.blob + (size_t)l * L.layer_stride;
        float* g = gp + (size_t)l * L.layer_stride;
        const float* sv = acts + l * per_layer;
        // ---- MLP: out = x_mid + gelu(fc_pre) Wpr + bpr ; fc_pre = a2 Wfc + bfc ; a2 = LN2(x_mid) ----
        tile_load(t1, sv + O_FC * TE, 4 * TE);
        __syncthreads();
        for (int i = tid; i < 4 * TE; i += blockDim.x) t2[i] = gelu_new(t1[i]);
        __syncthreads();
        mm_xTy(g + L.wpr, g + L.bpr, t2, h, T, 4 * E, E);
        mm_xwT<4>(t0, h, w + L.wpr, T, 4 * E, E, stage);                // d gelu_out [T][4E]
        __syncthreads();
        for (int i = tid; i < 4 * TE; i += blockDim.x) t0[i] *= gelu_new_grad(t1[i]);     // d fc_pre
        tile_load(t2, sv + O_A2 * TE, TE);
        __syncthreads();
        mm_xTy(g + L.wfc, g + L.bfc, t2, t0, T, E, 4 * E);
        mm_xwT<1>(t1, t0, w + L.wfc, T, E, 4 * E, stage);               // d a2 [T][E]
        tile_load(t3, sv + O_XMID * TE, TE);
        __syncthreads();
        ln_bwd(h, true, t3, t1, w + L.ln2w, g + L.ln2w, g + L.ln2b, T, E, p.eps);
        __syncthreads();
        // ---- attention block: x_mid = x_in + att Wproj + bproj ----
        tile_load(t0, sv + O_ATT * TE, TE);
        tile_load_ld(t2, LQ, sv + O_QKV * TE, T, 3 * E);
        __syncthreads();
        mm_xTy(g + L.wproj, g + L.bproj, t0, h, T, E, E);
        mm_xwT<1>(t1, h, w + L.wproj, T, E, E, stage);                  // d att [T][E]
        __syncthreads();
        // per head: P, dS in shared memory, then dq, dk, dv into t3 ([T][3E])
        const float rscale = 1.0f / sqrtf((float)hd);
        for (int hh = 0; hh < H; hh++) {
            for (int i = warp; i < T; i += nw) {
                float* Prow = P + (size_t)i * T;
                attn_row(t2, i, hh, T, E, hd, Prow, lane);
                // dP_ij = d_o_i . v_j ; rs_i = sum_j dP_ij P_ij ; dS_ij = P_ij (dP_ij - rs_i)
                const float* doi = t1 + i * E + hh * hd;
                float rs = 0.f;
                for (int j = lane; j <= i; j += 32) {
                    const float* v = t2 + j * LQ + 2 * E + hh * hd;
                    float dp = 0.f;
                    for (int d = 0; d < hd; d++) dp = fmaf(doi[d], v[d], dp);
                    dS[i * T + j] = dp;
                    rs = fmaf(dp, Prow[j], rs);
                }
                rs = warp_sum(rs);
                for (int j = lane; j < T; j += 32) dS[i * T + j] = (j <= i) ? Prow[j] * (dS[i * T + j] - rs) : 0.f;
            }
            __syncthreads();
            for (int idx = tid; idx < T * hd; idx += blockDim.x) {
                const int r = idx / hd, d = idx - r * hd;
                // dq_r = sum_{j<=r} dS_rj k_j / sqrt(hd)
                float dq = 0.f;
                for (int j = 0; j <= r; j++) dq = fmaf(dS[r * T + j], t2[j * LQ + E + hh * hd + d], dq);
                // dk_r = sum_{i>=r} dS_ir q_i / sqrt(hd) ; dv_r = sum_{i>=r} P_ir d_o_i
                float dk = 0.f, dv = 0.f;
                for (int i = r; i < T; i++) {
                    dk = fmaf(dS[i * T + r], t2[i * LQ + hh * hd + d], dk);
                    dv = fmaf(P[i * T + r], t1[i * E + hh * hd + d], dv);
                }
                t3[r * 3 * E + hh * hd + d] = dq * rscale;
                t3[r * 3 * E + E + hh * hd + d] = dk * rscale;
                t3[r * 3 * E + 2 * E + hh * hd + d] = dv;
            }
            __syncthreads();
        }
        tile_load(t0, sv + O_A1 * TE, TE);
        __syncthreads();
        mm_xTy(g + L.wqkv, g + L.bqkv, t0, t3, T, E, 3 * E);
        mm_xwT<1>(t1, t3, w + L.wqkv, T, E, 3 * E, stage);              // d a1 [T][E]
        tile_load(t2, sv + O_XIN * TE, TE);
        __syncthreads();
        ln_bwd(h, true, t2, t1, w + L.ln1w, g + L.ln1w, g + L.ln1b, T, E, p.eps);
        __syncthreads();
    }
}

// grad[i] = sum_b gpart[b][i] (fixed order) ; loss = sum_b loss_part[b] / (B*T*E)
__global__ void train_finalize_kernel(const float* __restrict__ gpart, const float* __restrict__ loss_part,
                                      float* __restrict__ grad, float* __restrict__ loss, int B, int total, float inv_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        float s = 0.f;
        for (int b = 0; b < B; b++) s += gpart[(size_t)b * total + i];
        grad[i] = s;
    }
    if (i == 0) {
        float s = 0.f;
        for (int b = 0; b < B; b++) s += loss_part[b];
        loss[0] = s * inv_count;
    }
}

}  // namespace

extern "C" {

int trpx_gpt2_train_fwd_bwd(trpx_gpt2_t h, const float* x, const float* labels, int B, int T, float* loss_dev,
                            float* hidden_dev, float* grad_blob_dev, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "weights not loaded");
    TRPX_REQUIRE(B >= 1 && T >= 1, "bad batch / length");
    TRPX_REQUIRE(T <= h->cfg.n_ctx, "sequence length %d exceeds n_ctx %d", T, h->cfg.n_ctx);
    TRPX_REQUIRE(x != nullptr && labels != nullptr && loss_dev != nullptr && grad_blob_dev != nullptr, "null pointer");
    DeviceGuard guard(h->device);
    const Gpt2Layout& L = h->lay;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t TE = (size_t)T * L.E;
    // E = 32 / 4 heads / T <= 64: every contraction on tensor cores (gpt2_train_tc.cu); TRPX_TRAIN_SIMT=1 keeps the FP32 kernel
    const char* simt_env = getenv("TRPX_TRAIN_SIMT");
    if (train_tc_supported(L, T) && !(simt_env && simt_env[0] == '1')) {
        const size_t aps = train_tc_acts_per_seq(L, T);
        const size_t need_tc = sizeof(float) * ((size_t)B * (aps + L.total + 1));
        if (need_tc > h->train_ws_bytes) {
            if (h->train_ws) TRPX_CUDA(cudaFree(h->train_ws));
            h->train_ws = nullptr;
            h->train_ws_bytes = 0;
            TRPX_CUDA(cudaMalloc(&h->train_ws, need_tc));
            h->train_ws_bytes = need_tc;
        }
        float* gpart = h->train_ws + (size_t)B * aps;
        float* loss_part = gpart + (size_t)B * L.total;
        int rc = launch_train_tc(L, h->blob, h->pe, x, labels, hidden_dev, h->train_ws, gpart, loss_part, B, T, h->cfg.ln_eps, st);
        if (rc) return rc;
        train_finalize_kernel<<<(L.total + 255) / 256, 256, 0, st>>>(gpart, loss_part, grad_blob_dev, loss_dev, B, L.total,
                                                                      1.0f / ((float)B * (float)T * (float)L.E));
        TRPX_CUDA(cudaGetLastError());
        return TRPX_OK;
    }
    TRPX_REQUIRE(TE <= 2048, "training step supports T * n_embd <= 2048");   // unit counts of mm_xwT<MAXU>
    TRPX_REQUIRE((L.hd & (L.hd - 1)) == 0, "training step supports power-of-two head sizes");
    const size_t smem = sizeof(float) * (TE * 17 + ((size_t)std::max(T, TR_THREADS / 32) + T) * T + STAGE_FLOATS);
    TRPX_REQUIRE(smem + 1024 <= (size_t)h->smem_optin, "training step needs %zu B of shared memory per sequence", smem);
    const size_t acts_per_seq = (size_t)L.n_layer * 12 * TE + 2 * TE;
    const size_t need = sizeof(float) * ((size_t)B * (acts_per_seq + L.total + 1));
    if (need > h->train_ws_bytes) {
        if (h->train_ws) TRPX_CUDA(cudaFree(h->train_ws));
        h->train_ws = nullptr;
        h->train_ws_bytes = 0;
        TRPX_CUDA(cudaMalloc(&h->train_ws, need));
        h->train_ws_bytes = need;
    }
    TrainParams p{};
    p.blob = h->blob; p.pe = h->pe; p.x = x; p.labels = labels; p.hidden = hidden_dev;
    p.acts = h->train_ws;
    p.gpart = h->train_ws + (size_t)B * acts_per_seq;
    p.loss_part = p.gpart + (size_t)B * L.total;
    p.B = B; p.T = T; p.eps = h->cfg.ln_eps;
    p.dscale = 2.0f / ((float)B * (float)T * (float)L.E);
    p.lay = L;
    TRPX_CUDA(cudaFuncSetAttribute(train_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    train_fwd_bwd_kernel<<<B, TR_THREADS, smem, st>>>(p);
    TRPX_CUDA(cudaGetLastError());
    train_finalize_kernel<<<(L.total + 255) / 256, 256, 0, st>>>(p.gpart, p.loss_part, grad_blob_dev, loss_dev, B, L.total,
                                                                  1.0f / ((float)B * (float)T * (float)L.E));
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

int trpx_gpt2_blob_size(trpx_gpt2_t h) { return valid(h) ? h->lay.total : -1; }

}  // extern "C"
