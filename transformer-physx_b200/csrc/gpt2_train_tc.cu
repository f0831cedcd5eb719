// K8t: teacher-forced training step of the decoder stack on tensor cores, for the E = 32 / 4-head models (Lorenz, Rossler)
// with T <= 64 tokens per sequence.  Same contract as train_fwd_bwd_kernel (gpt2_train.cu):
//     hidden = transformer.forward(x)[0];  loss = MSELoss()(hidden, labels)            phys_transformer_helpers.py:36-69
// and d loss / d parameter for the whole stack in the packed blob layout (utils/trainer.py:244-277 does loss.backward()).
//
// ncu on the FP32 kernel showed its three tile products at only 35 % of the time and the per-head attention loops at 40 %,
// so here EVERY contraction is an error-compensated 3xTF32 mma.sync.m16n8k8 tile product on shared-memory tiles:
// the Conv1D / Linear products and their two backward products (dX = dY W^T, dW = X^T dY), and per head
// S = Q K^T, O = P V, dP = dO V^T, dV = P^T dO, dQ = dS K, dK = dS^T Q.  One generic routine (tc_gemm) takes both operands
// through (row stride, column stride) so transposed operands need no staging.  A CTA (8 warps) owns one sequence as a
// 64-row tile (rows >= T are zero and stay zero through the backward); 112 KB of shared memory => two sequences per SM.
// The forward saves x_in, q|k|v, att, x_mid and fc_pre per layer (L2-resident workspace); LayerNorm outputs, gelu and the
// attention probabilities are recomputed in the backward.  Weight gradients leave as per-sequence partials and are summed
// in a fixed order by train_finalize_kernel (deterministic, no atomics).
#include "common.cuh"
#include "gpt2_layout.cuh"
#include "gpt2_handle.cuh"
#include "gpt2_train_tc.cuh"
#include "tc_tile.cuh"

namespace trpx {
namespace {

using namespace tct;                       // TT = 256 threads (8 warps), tc_gemm
constexpr int TM = 64;                     // rows of the token tile
constexpr int LDX = 36, LDQ = 100, LDH = 132, LDP = 68, PHS = 64 * LDP + 8;      // PHS: per-head stride of the P tiles
constexpr int WORK = 17472;                // floats of the work region: two [64][132] MLP tiles, or 4 score tiles, or the attention-backward set
constexpr int SM_FLOATS = 2 * TM * LDX + TM * LDQ + WORK;

// gelu_new and its derivative in sigmoid form (0.5 x (1 + tanh u) == x sigma(2u), u = c (x + 0.044715 x^3);
// transformer/utils.py:57-60): one __expf + one division for both, relative deviation ~1e-7.
__device__ __forceinline__ void gelu_and_grad(float x, float& gval, float& grad) {
    const float c = 0.7978845608028654f;
    const float x2 = x * x;
    const float u2 = 2.0f * c * x * (1.0f + 0.044715f * x2);
    const float sg = __fdividef(1.0f, 1.0f + __expf(-u2));
    gval = x * sg;
    grad = sg + x * sg * (1.0f - sg) * (2.0f * c) * (1.0f + 3.0f * 0.044715f * x2);
}

// LayerNorm over the 64 rows of a tile, all rows at once: thread = (row tid >> 2, quarter tid & 3) owns 8 channels
// (two float4), the row statistics take two xor-shuffles over the 4 lanes of the row.
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(FULL, v, 1);
    v += __shfl_xor_sync(FULL, v, 2);
    return v;
}
// Y = LayerNorm(X) (row strides ldx / ldy, multiples of 4)
__device__ __forceinline__ void ln_fwd_tile(const float* X, int ldx, float* Y, int ldy, const float* __restrict__ w,
                                            const float* __restrict__ b, float eps) {
    const int r = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 8;
    float x[8], wv[8], bv[8];
    ld8(X + r * ldx + c0, x);
    ld8(w + c0, wv);
    ld8(b + c0, bv);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    const float mean = quad_sum(s) * (1.0f / 32.0f);
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] -= mean; v = fmaf(x[i], x[i], v); }
    const float rstd = 1.0f / sqrtf(quad_sum(v) * (1.0f / 32.0f) + eps);
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = x[i] * rstd * wv[i] + bv[i];
    st8(Y + r * ldy + c0, x);
}
// LayerNorm backward: Xin = LN input (ld LDX), dY = gradient of the LN output; dH[r][c] (+)= dx.  The per-row products
// dY * xhat and dY are left in `tmp` ([64][72]: 32 + 32 columns, row stride 72) and column-summed by 64 threads into the
// weight / bias gradients gw / gb.  The caller synchronises before (inputs ready) and after (dH visible, tmp reusable).
constexpr int LDT = 72;
__device__ __forceinline__ void ln_bwd_tile(float* dH, bool accumulate, const float* Xin, const float* dY, int ldy,
                                            const float* __restrict__ w, float* gw, float* gb, float eps, float* tmp) {
    const int r = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 8;
    float x[8], dy[8], wv[8];
    ld8(Xin + r * LDX + c0, x);
    ld8(dY + r * ldy + c0, dy);
    ld8(w + c0, wv);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    const float mean = quad_sum(s) * (1.0f / 32.0f);
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] -= mean; v = fmaf(x[i], x[i], v); }
    const float rstd = 1.0f / sqrtf(quad_sum(v) * (1.0f / 32.0f) + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x[i] *= rstd;                                   // xhat
        const float dyw = dy[i] * wv[i];
        m1 += dyw;
        m2 = fmaf(dyw, x[i], m2);
    }
    m1 = quad_sum(m1) * (1.0f / 32.0f);
    m2 = quad_sum(m2) * (1.0f / 32.0f);
    float dx[8], pw[8];
    if (accumulate) ld8(dH + r * LDX + c0, dx);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float d = rstd * (dy[i] * wv[i] - m1 - x[i] * m2);
        dx[i] = accumulate ? dx[i] + d : d;
        pw[i] = dy[i] * x[i];
    }
    st8(dH + r * LDX + c0, dx);
    st8(tmp + r * LDT + c0, pw);
    st8(tmp + r * LDT + 32 + c0, dy);
    __syncthreads();
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int t = 0; t < TM; t += 4) {
            a0 += tmp[t * LDT + c];
            a1 += tmp[(t + 1) * LDT + c];
            a2 += tmp[(t + 2) * LDT + c];
            a3 += tmp[(t + 3) * LDT + c];
        }
        (c < 32 ? gw : gb)[c & 31] = (a0 + a1) + (a2 + a3);
    }
}
// out[n] = sum over the 64 rows of dY[r][n]
__device__ __forceinline__ void colsum(const float* dY, int ld, int N, float* out) {
    for (int n = threadIdx.x; n < N; n += TT) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int r = 0; r < TM; r += 4) {
            a0 += dY[r * ld + n];
            a1 += dY[(r + 1) * ld + n];
            a2 += dY[(r + 2) * ld + n];
            a3 += dY[(r + 3) * ld + n];
        }
        out[n] = (a0 + a1) + (a2 + a3);
    }
}
// tile <-> global: [T][n] contiguous rows in global (n % 4 == 0, 16-byte aligned), padded rows in shared memory.
// Fetch: 16-byte cp.async copies, all in flight at once (a plain load loop exposed one L2/HBM latency per iteration);
// rows >= T of the tile are zero-filled.  The caller's __syncthreads() publishes the tile.
__device__ __forceinline__ void tile_save(float* __restrict__ dst, const float* src, int ld, int T, int n) {
    const int nq = n >> 2;
    for (int i = threadIdx.x; i < T * nq; i += TT) {
        const int r = i / nq, c = (i - r * nq) << 2;
        *reinterpret_cast<float4*>(dst + (size_t)r * n + c) = *reinterpret_cast<const float4*>(src + r * ld + c);
    }
}
__device__ __forceinline__ void tile_fetch(float* dst, int ld, const float* __restrict__ src, int T, int n) {
    const int nq = n >> 2;
    for (int i = threadIdx.x; i < TM * nq; i += TT) {
        const int r = i / nq, c = (i - r * nq) << 2;
        float* d = dst + r * ld + c;
        if (r < T) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d)), "l"(src + (size_t)r * n + c) : "memory");
        } else {
            *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}
__device__ __forceinline__ void tile_fetch_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Pull a weight matrix into L1 ahead of the tile product that reads it (fire and forget, one 128-byte line per thread):
// with 2 x 111 KB of shared memory per SM only ~28 KB of L1 are left, so weights otherwise come from L2 on every use.
__device__ __forceinline__ void prefetch_l1(const float* p, int nfloats) {
    for (int i = threadIdx.x * 32; i < nfloats; i += TT * 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + i));
}

struct TcParams {
    const float* blob;
    const float* pe;
    const float* x;
    const float* labels;
    float* hidden;
    float* acts;
    float* gpart;
    float* loss_part;
    int B, T;
    float eps, dscale;
    Gpt2Layout lay;
};

// causal softmax of the rows of nh score tiles S[h][i][j] (log2 domain): 4 lanes per (row, head), no branch around shuffles
__device__ __forceinline__ void softmax_rows(float* P, int nh, int T) {
    for (int idx = threadIdx.x >> 2; idx < TM * nh; idx += TT / 4) {
        const int part = threadIdx.x & 3;
        const int hh = idx % nh, i = idx / nh;
        float* row = P + hh * PHS + i * LDP;
        const int lim = (i < T) ? i : -1;
        float m = -INFINITY;
        for (int j = part; j <= lim; j += 4) m = fmaxf(m, row[j]);
        m = fmaxf(m, __shfl_xor_sync(FULL, m, 1));
        m = fmaxf(m, __shfl_xor_sync(FULL, m, 2));
        float sum = 0.f;
        for (int j = part; j <= lim; j += 4) {
            const float e = ex2_approx(row[j] - m);
            row[j] = e;
            sum += e;
        }
        sum = quad_sum(sum);
        const float inv = (lim >= 0) ? 1.0f / sum : 0.f;
        for (int j = part; j < TM; j += 4) row[j] = (j <= lim) ? row[j] * inv : 0.f;
    }
}

__global__ void __launch_bounds__(TT, 2) train_e32_tc_kernel(TcParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float lred[8];
    const Gpt2Layout& L = p.lay;
    constexpr int E = 32, H = 4, HD = 8;
    const int T = p.T, NL = L.n_layer;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* X = sm;                           // [64][36]  residual stream (forward) / its gradient (backward)
    float* A = X + TM * LDX;                 // [64][36]  GEMM input / LayerNorm input in the backward
    float* QKV = A + TM * LDX;               // [64][100] q | k | v  (also: LayerNorm outputs [64][36] in the backward)
    float* Wk = QKV + TM * LDQ;              // work region: MLP tiles [64][132] x 2, score tiles, gradient tiles
    float* H1 = Wk;                          // [64][132]
    float* H2 = Wk + TM * LDH;               // [64][132]
    const int b = blockIdx.x;
    const int TE = T * E;
    // saved per layer: x_in [TE] | qkv [3TE] | att [TE] | x_mid [TE] | fc_pre [4TE]; then h_L [TE]
    const size_t per_layer = (size_t)10 * TE;
    float* acts = p.acts + (size_t)b * (NL * per_layer + (size_t)TE);
    float* gp = p.gpart + (size_t)b * L.total;
    const float qscale = 1.4426950408889634f * 0.35355339059327373f;       // log2(e) / sqrt(hd), attention.py:97-99

    // ============================== forward ==============================
    for (int i = tid; i < TM * E; i += TT) {
        const int t = i >> 5, c = i & 31;
        X[t * LDX + c] = (t < T) ? p.x[(size_t)b * TE + i] + __ldg(p.pe + t * E + c) : 0.f;   // position id = t
    }
    __syncthreads();
    for (int l = 0; l < NL; l++) {
        const float* w = p.blob + (size_t)l * L.layer_stride;
        float* sv = acts + l * per_layer;
        prefetch_l1(w + L.wqkv, 3 * E * E);
        tile_save(sv, X, LDX, T, E);
        ln_fwd_tile(X, LDX, A, LDX, w + L.ln1w, w + L.ln1b, p.eps);
        __syncthreads();
        {
            const float* bias = w + L.bqkv;
            tc_gemm<2, 3, 4, true, false>(A, 0, LDX, 1, w + L.wqkv, 0, 3 * E, 1, 1, 4, 12,[&](int, int m, int n, float v) { QKV[m * LDQ + n] = v + __ldg(bias + n); });
        }
        __syncthreads();
        tile_save(sv + TE, QKV, LDQ, T, 3 * E);
        prefetch_l1(w + L.wproj, E * E);
        // scores of the 4 heads at once: S_h = Q_h K_h^T (log2 domain) into the work region
        tc_gemm<2, 4, 1, true, true>(QKV, HD, LDQ, 1, QKV + E, HD, 1, LDQ, H, 4, 8,[&](int hh, int m, int n, float v) { Wk[hh * PHS + m * LDP + n] = v * qscale; });
        __syncthreads();
        softmax_rows(Wk, H, T);
        __syncthreads();
        tc_gemm<2, 1, 8, true, false>(Wk, PHS, LDP, 1, QKV + 2 * E, HD, LDQ, 1, H, 4, 1,[&](int hh, int m, int n, float v) { A[m * LDX + hh * HD + n] = v; });
        __syncthreads();
        tile_save(sv + 4 * TE, A, LDX, T, E);
        prefetch_l1(w + L.wfc, 4 * E * E);
        {
            const float* bias = w + L.bproj;
            tc_gemm<1, 2, 4, true, false>(A, 0, LDX, 1, w + L.wproj, 0, E, 1, 1, 4, 4,[&](int, int m, int n, float v) { X[m * LDX + n] += v + __ldg(bias + n); });
        }
        __syncthreads();
        tile_save(sv + 5 * TE, X, LDX, T, E);
        prefetch_l1(w + L.wpr, 4 * E * E);
        ln_fwd_tile(X, LDX, A, LDX, w + L.ln2w, w + L.ln2b, p.eps);
        __syncthreads();
        {
            const float* bias = w + L.bfc;
            float* fsave = sv + 6 * TE;
            tc_gemm<2, 4, 4, true, false>(A, 0, LDX, 1, w + L.wfc, 0, 4 * E, 1, 1, 4, 16,[&](int, int m, int n, float v) {
                const float f = v + __ldg(bias + n);
                if (m < T) fsave[(size_t)m * 4 * E + n] = f;
                H2[m * LDH + n] = gelu_new_sig(f);
            });
        }
        __syncthreads();
        {
            const float* bias = w + L.bpr;
            tc_gemm<1, 2, 16, true, false>(H2, 0, LDH, 1, w + L.wpr, 0, E, 1, 1, 4, 4,[&](int, int m, int n, float v) { X[m * LDX + n] += v + __ldg(bias + n); });
        }
        __syncthreads();
    }
    // padded rows picked up the biases: clear them so that they stay out of every gradient
    for (int i = tid; i < (TM - T) * E; i += TT) X[(T + (i >> 5)) * LDX + (i & 31)] = 0.f;
    __syncthreads();
    tile_save(acts + NL * per_layer, X, LDX, T, E);
    ln_fwd_tile(X, LDX, A, LDX, p.blob + L.lnfw, p.blob + L.lnfb, p.eps);
    __syncthreads();
    // head + loss: dY (ld 36) into H1
    float lsum = 0.f;
    {
        const float* bias = p.blob + L.bf;
        tc_gemm<1, 2, 4, true, false>(A, 0, LDX, 1, p.blob + L.wf, 0, E, 1, 1, 4, 4,[&](int, int m, int n, float v) {
            float d = 0.f;
            if (m < T) {
                const float y = v + __ldg(bias + n);
                const size_t o = (size_t)b * TE + (size_t)m * E + n;
                if (p.hidden) p.hidden[o] = y;
                d = y - p.labels[o];
                lsum = fmaf(d, d, lsum);
            }
            H1[m * LDX + n] = d * p.dscale;
        });
    }
    lsum = warp_sum(lsum);
    if (lane == 0) lred[warp] = lsum;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; i++) s += lred[i];
        p.loss_part[b] = s;
    }

    // ============================== backward ==============================
    // head: y = a_f Wf + bf (blob wf is [in][out]); a_f = LN_f(h_L) is in A, dY in H1 (ld 36), h_L in X
    {
        float* gwf = gp + L.wf;
        tc_gemm<1, 1, 8, false, false>(A, 0, 1, LDX, H1, 0, LDX, 1, 1, 2, 4,[&](int, int m, int n, float v) { gwf[m * E + n] = v; });
        colsum(H1, LDX, E, gp + L.bf);
        tc_gemm<1, 2, 4, true, true>(H1, 0, LDX, 1, p.blob + L.wf, 0, 1, E, 1, 4, 4,[&](int, int m, int n, float v) { H2[m * LDX + n] = v; });
    }
    __syncthreads();
    // X (h_L) -> dh in place: ln_bwd reads Xin and writes dH row by row, each lane its own element
    ln_bwd_tile(X, false, X, H2, LDX, p.blob + L.lnfw, gp + L.lnfw, gp + L.lnfb, p.eps, H1);          // dY of the head is dead
    __syncthreads();
    for (int l = NL - 1; l >= 0; l--) {
        const float* w = p.blob + (size_t)l * L.layer_stride;
        float* g = gp + (size_t)l * L.layer_stride;
        const float* sv = acts + l * per_layer;
        float* LNO = QKV;                                  // [64][36] LayerNorm output of this phase
        // ---- MLP: out = x_mid + gelu(fc_pre) Wpr + bpr ; fc_pre = a2 Wfc + bfc ; a2 = LN2(x_mid) ----
        prefetch_l1(w + L.wpr, 4 * E * E);
        tile_fetch(H1, LDH, sv + 6 * TE, T, 4 * E);       // fc_pre
        tile_fetch(A, LDX, sv + 5 * TE, T, E);            // x_mid
        tile_fetch_wait();
        __syncthreads();
        for (int i = tid; i < TM * 4 * E; i += TT) {
            const int r = i >> 7, c = i & 127;
            float gv, gd;
            gelu_and_grad(H1[r * LDH + c], gv, gd);
            H2[r * LDH + c] = gv;                          // gelu(fc_pre)
            H1[r * LDH + c] = gd;                          // gelu'(fc_pre), consumed by the d fc_pre epilogue
        }
        ln_fwd_tile(A, LDX, LNO, LDX, w + L.ln2w, w + L.ln2b, p.eps);
        __syncthreads();
        {
            prefetch_l1(w + L.wfc, 4 * E * E);
            float* gw = g + L.wpr;                          // dWpr = gelu_out^T dh   [128][32]
            tc_gemm<2, 2, 8, false, false>(H2, 0, 1, LDH, X, 0, LDX, 1, 1, 8, 4,[&](int, int m, int n, float v) { gw[m * E + n] = v; });
            colsum(X, LDX, E, g + L.bpr);
        }
        __syncthreads();
        // d fc_pre = (dh Wpr^T) * gelu'(fc_pre) -> H2
        tc_gemm<2, 4, 4, true, true>(X, 0, LDX, 1, w + L.wpr, 0, 1, E, 1, 4, 16,[&](int, int m, int n, float v) { H2[m * LDH + n] = v * H1[m * LDH + n]; });
        __syncthreads();
        {
            float* gw = g + L.wfc;                          // dWfc = a2^T d_fc   [32][128]
            tc_gemm<2, 2, 8, false, false>(LNO, 0, 1, LDX, H2, 0, LDH, 1, 1, 2, 16,[&](int, int m, int n, float v) { gw[m * 4 * E + n] = v; });
            colsum(H2, LDH, 4 * E, g + L.bfc);
            // d a2 = d_fc Wfc^T -> H1 (ld 36); fc_pre is no longer needed
        }
        __syncthreads();
        tc_gemm<1, 2, 16, true, true>(H2, 0, LDH, 1, w + L.wfc, 0, 1, 4 * E, 1, 4, 4,[&](int, int m, int n, float v) { H1[m * LDX + n] = v; });
        __syncthreads();
        ln_bwd_tile(X, true, A, H1, LDX, w + L.ln2w, g + L.ln2w, g + L.ln2b, p.eps, H2);                   // d_fc is dead
        __syncthreads();
        // ---- attention block: x_mid = x_in + att Wproj + bproj ----
        prefetch_l1(w + L.wproj, E * E);
        tile_fetch(A, LDX, sv + 4 * TE, T, E);            // att
        tile_fetch(QKV, LDQ, sv + TE, T, 3 * E);          // q | k | v
        tile_fetch_wait();
        __syncthreads();
        float* dAtt = Wk;                                  // [64][36]
        float* D3 = Wk + TM * LDX;                         // [64][100] d(q | k | v)
        float* Ph = D3 + TM * LDQ;                         // [64][65]  probabilities of the current head
        float* dSh = Ph + TM * LDP;                        // [64][65]  dP, then dS
        {
            float* gw = g + L.wproj;                        // dWproj = att^T dh
            tc_gemm<1, 1, 8, false, false>(A, 0, 1, LDX, X, 0, LDX, 1, 1, 2, 4,[&](int, int m, int n, float v) { gw[m * E + n] = v; });
            colsum(X, LDX, E, g + L.bproj);
            tc_gemm<1, 2, 4, true, true>(X, 0, LDX, 1, w + L.wproj, 0, 1, E, 1, 4, 4,[&](int, int m, int n, float v) { dAtt[m * LDX + n] = v; });
        }
        __syncthreads();
        for (int hh = 0; hh < H; hh++) {
            const float* Qh = QKV + hh * HD;
            const float* Kh = QKV + E + hh * HD;
            const float* Vh = QKV + 2 * E + hh * HD;
            const float* dOh = dAtt + hh * HD;
            // S = Q K^T (log2 domain) -> Ph ; dP = dO V^T -> dSh
            tc_gemm<2, 4, 1, true, true>(Qh, 0, LDQ, 1, Kh, 0, 1, LDQ, 1, 4, 8,[&](int, int m, int n, float v) { Ph[m * LDP + n] = v * qscale; });
            tc_gemm<2, 4, 1, true, true>(dOh, 0, LDX, 1, Vh, 0, 1, LDQ, 1, 4, 8,[&](int, int m, int n, float v) { dSh[m * LDP + n] = v; }, 4);
            __syncthreads();
            // rows: P = softmax(S) (causal), dS = P (dP - sum_j P dP) / sqrt(hd); 4 lanes per row (rows >= T: all zero;
            // no branch around the shuffles)
            {
                const int i = tid >> 2, part = tid & 3;
                float* pr = Ph + i * LDP;
                float* dr = dSh + i * LDP;
                const int lim = (i < T) ? i : -1;                  // last valid key
                float m = -INFINITY;
                for (int j = part; j <= lim; j += 4) m = fmaxf(m, pr[j]);
                m = fmaxf(m, __shfl_xor_sync(FULL, m, 1));
                m = fmaxf(m, __shfl_xor_sync(FULL, m, 2));
                float sum = 0.f;
                for (int j = part; j <= lim; j += 4) {
                    const float e = ex2_approx(pr[j] - m);
                    pr[j] = e;
                    sum += e;
                }
                sum += __shfl_xor_sync(FULL, sum, 1);
                sum += __shfl_xor_sync(FULL, sum, 2);
                const float inv = (lim >= 0) ? 1.0f / sum : 0.f;
                float rs = 0.f;
                for (int j = part; j <= lim; j += 4) {
                    const float pv = pr[j] * inv;
                    pr[j] = pv;
                    rs = fmaf(pv, dr[j], rs);
                }
                rs += __shfl_xor_sync(FULL, rs, 1);
                rs += __shfl_xor_sync(FULL, rs, 2);
                for (int j = part; j < TM; j += 4) {
                    if (j <= lim) {
                        dr[j] = pr[j] * (dr[j] - rs) * 0.35355339059327373f;
                    } else {
                        pr[j] = 0.f;
                        dr[j] = 0.f;
                    }
                }
            }
            __syncthreads();
            // dQ = dS K ; dK = dS^T Q ; dV = P^T dO
            tc_gemm<1, 1, 8, true, false>(dSh, 0, LDP, 1, Kh, 0, LDQ, 1, 1, 4, 1,[&](int, int m, int n, float v) { D3[m * LDQ + hh * HD + n] = v; });
            tc_gemm<1, 1, 8, false, false>(dSh, 0, 1, LDP, Qh, 0, LDQ, 1, 1, 4, 1,[&](int, int m, int n, float v) { D3[m * LDQ + E + hh * HD + n] = v; }, 4);
            tc_gemm<1, 1, 8, false, false>(Ph, 0, 1, LDP, dOh, 0, LDX, 1, 1, 4, 1,[&](int, int m, int n, float v) { D3[m * LDQ + 2 * E + hh * HD + n] = v; });
            __syncthreads();
        }
        // ---- q | k | v = a1 Wqkv + bqkv ; a1 = LN1(x_in) ----
        prefetch_l1(w + L.wqkv, 3 * E * E);
        tile_fetch(A, LDX, sv, T, E);                      // x_in
        tile_fetch_wait();
        __syncthreads();
        ln_fwd_tile(A, LDX, LNO, LDX, w + L.ln1w, w + L.ln1b, p.eps);      // q | k | v are dead: LNO aliases them
        __syncthreads();
        {
            float* gw = g + L.wqkv;                         // dWqkv = a1^T d(qkv)   [32][96]
            tc_gemm<1, 3, 8, false, false>(LNO, 0, 1, LDX, D3, 0, LDQ, 1, 1, 2, 12,[&](int, int m, int n, float v) { gw[m * 3 * E + n] = v; });
            colsum(D3, LDQ, 3 * E, g + L.bqkv);
            float* dA1 = Ph;                                // [64][36] (the score tiles are dead)
            tc_gemm<1, 2, 12, true, true>(D3, 0, LDQ, 1, w + L.wqkv, 0, 1, 3 * E, 1, 4, 4,[&](int, int m, int n, float v) { dA1[m * LDX + n] = v; });
        }
        __syncthreads();
        ln_bwd_tile(X, true, A, Ph, LDX, w + L.ln1w, g + L.ln1w, g + L.ln1b, p.eps, D3);                   // d(qkv) is dead
        __syncthreads();
    }
}

}  // namespace

bool train_tc_supported(const Gpt2Layout& L, int T) { return L.E == 32 && L.n_head == 4 && T >= 1 && T <= TM; }
size_t train_tc_acts_per_seq(const Gpt2Layout& L, int T) { return (size_t)L.n_layer * 10 * T * L.E + (size_t)T * L.E; }

int launch_train_tc(const Gpt2Layout& L, const float* blob, const float* pe, const float* x, const float* labels, float* hidden,
                    float* acts, float* gpart, float* loss_part, int B, int T, float eps, cudaStream_t st) {
    TcParams p{};
    p.blob = blob; p.pe = pe; p.x = x; p.labels = labels; p.hidden = hidden; p.acts = acts; p.gpart = gpart;
    p.loss_part = loss_part; p.B = B; p.T = T; p.eps = eps;
    p.dscale = 2.0f / ((float)B * (float)T * (float)L.E);
    p.lay = L;
    const size_t smem = SM_FLOATS * sizeof(float);
    TRPX_CUDA(cudaFuncSetAttribute(train_e32_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    train_e32_tc_kernel<<<B, TT, smem, st>>>(p);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

}  // namespace trpx
