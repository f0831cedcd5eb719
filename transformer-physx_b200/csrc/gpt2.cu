// Decoder stack of the rollout path: weight packer, rollout kernels (K1), teacher-forced forward (K2).
//
// Reference arithmetic restated once (validated by oracle/trpx_oracle.py):
//   h = x + PE(p);  per layer:  a = LN1(h); [q|k|v] = a Wqkv + b;  window <- (k,v);
//   per head: w = softmax(q K^T / sqrt(hd)); o = w V;  h += o Wproj + b;
//   f = LN2(h);  h += gelu_new(f Wfc + b) Wpr + b;      x' = LN_f(h) Wf^T + bf
// generate(): trphysx/transformer/generate_utils.py:112-169 ; forward(): phys_transformer_gpt2.py:147-269 ;
// attention: attention.py:75-119, 153-200.
#include "common.cuh"
#include "gpt2_layout.cuh"
#include "gpt2_handle.cuh"
#include "gpt2_tmem.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

using namespace trpx;

namespace {

constexpr int GEN_THREADS = 256;      // forward kernel
constexpr int ROLL_THREADS = 512;     // generic rollout kernel

// ------------------------------------------------------------------------------------------------
// building blocks shared by the generic rollout kernel and the forward kernel (CTA-cooperative)
// ------------------------------------------------------------------------------------------------

// y[g][n] = bias[n] + sum_k aT[k*G+g] * W[k*N+n]   (W row-major [K][N] in global memory, read once per CTA)
// Threads own output columns; when N < blockDim the K range is split across thread groups and
// combined through `part` (>= blockDim*G floats).  epi(n, g, value) consumes the result.
// All threads of the CTA must call this; the caller synchronises afterwards.
template <int G, class Epi>
__device__ __forceinline__ void block_gemv(const float* __restrict__ W, const float* __restrict__ bias, int K,
                                           int N, const float* aT, float* part, Epi epi) {
    const int NT = blockDim.x, tid = threadIdx.x;
    if (N >= NT) {
        for (int n = tid; n < N; n += NT) {
            float acc[G];
            const float b = __ldg(bias + n);
#pragma unroll
            for (int g = 0; g < G; g++) acc[g] = b;
#pragma unroll 4
            for (int k = 0; k < K; k++) {
                const float w = __ldg(W + (size_t)k * N + n);
#pragma unroll
                for (int g = 0; g < G; g++) acc[g] = fmaf(aT[k * G + g], w, acc[g]);
            }
#pragma unroll
            for (int g = 0; g < G; g++) epi(n, g, acc[g]);
        }
    } else {
        const int ks = NT / N;
        const int kc = (K + ks - 1) / ks;
        const int my = tid / N, n = tid - my * N;
        if (my < ks) {
            float acc[G];
#pragma unroll
            for (int g = 0; g < G; g++) acc[g] = 0.f;
            const int k0 = my * kc, k1 = min(K, k0 + kc);
#pragma unroll 4
            for (int k = k0; k < k1; k++) {
                const float w = __ldg(W + (size_t)k * N + n);
#pragma unroll
                for (int g = 0; g < G; g++) acc[g] = fmaf(aT[k * G + g], w, acc[g]);
            }
#pragma unroll
            for (int g = 0; g < G; g++) part[(my * N + n) * G + g] = acc[g];
        }
        __syncthreads();
        for (int idx = tid; idx < N * G; idx += NT) {
            const int nn = idx / G, g = idx - nn * G;
            float s = __ldg(bias + nn);
            for (int j = 0; j < ks; j++) s += part[(j * N + nn) * G + g];
            epi(nn, g, s);
        }
    }
}

// Same contract as block_gemv, but every thread owns FOUR adjacent output columns and streams the weights with
// 128-bit loads (4x fewer load instructions, 4x the bytes in flight), with the K range split over
// blockDim / (N/4) thread groups.  `part` must hold 4 * blockDim * G floats.  Used where the weights stream from
// L2 every step (Cylinder: 4.6 MB per step per CTA), which is what bounds the generic kernel.
template <int G, class Epi>
__device__ __forceinline__ void block_gemv4(const float* __restrict__ W, const float* __restrict__ bias, int K,
                                            int N, const float* aT, float* part, Epi epi) {
    const int NT = blockDim.x, tid = threadIdx.x;
    const int NQ = N >> 2;
    if (NQ > NT) {                                  // very wide layer: fall back to the scalar-column version
        block_gemv<G>(W, bias, K, N, aT, part, epi);
        return;
    }
    const int ks = NT / NQ;
    const int kc = (K + ks - 1) / ks;
    const int my = tid / NQ, nq = tid - my * NQ;
    if (my < ks) {
        float acc[G][4];
#pragma unroll
        for (int g = 0; g < G; g++) acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.f;
        const int k0 = my * kc, k1 = min(K, k0 + kc);
        const float4* wp = reinterpret_cast<const float4*>(W) + nq;
#pragma unroll 8
        for (int k = k0; k < k1; k++) {
            const float4 w = __ldg(wp + (size_t)k * NQ);
#pragma unroll
            for (int g = 0; g < G; g++) {
                const float a = aT[k * G + g];
                acc[g][0] = fmaf(a, w.x, acc[g][0]);
                acc[g][1] = fmaf(a, w.y, acc[g][1]);
                acc[g][2] = fmaf(a, w.z, acc[g][2]);
                acc[g][3] = fmaf(a, w.w, acc[g][3]);
            }
        }
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
            for (int j = 0; j < 4; j++) part[((size_t)my * N + nq * 4 + j) * G + g] = acc[g][j];
    }
    __syncthreads();
    for (int idx = tid; idx < N * G; idx += NT) {
        const int nn = idx / G, g = idx - nn * G;
        float s2 = __ldg(bias + nn);
        for (int j = 0; j < ks; j++) s2 += part[((size_t)j * N + nn) * G + g];
        epi(nn, g, s2);
    }
}

// One warp: softmax(q K^T / sqrt(hd)) V for one (query, head).  K/V rows have `E` floats, this head's
// slice starts at column head*hd.  q element d is q[d*qs]; output element d goes to out[d*os].
// pbuf: per-warp scratch of >= L floats.  If attn_out != nullptr the L weights are stored there.
__device__ __forceinline__ void warp_attention(const float* q, int qs, const float* Kr, const float* Vr, int E,
                                               int hd, int L, float* pbuf, float* out, int os, float* attn_out) {
    const int lane = threadIdx.x & 31;
    const float scale = sqrtf((float)hd);          // attention.py:99  w / (float(v.size(-1)) ** 0.5)
    float m = -INFINITY;
    for (int slot = lane; slot < L; slot += 32) {
        const float* kr = Kr + (size_t)slot * E;
        float s = 0.f;
        for (int d = 0; d < hd; d++) s = fmaf(q[d * qs], kr[d], s);
        s = s / scale;
        pbuf[slot] = s;
        m = fmaxf(m, s);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int slot = lane; slot < L; slot += 32) {
        const float p = expf(pbuf[slot] - m);
        pbuf[slot] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    for (int slot = lane; slot < L; slot += 32) {
        const float p = pbuf[slot] / sum;
        pbuf[slot] = p;
        if (attn_out) attn_out[slot] = p;
    }
    __syncwarp();
    const int ld = hd < 32 ? hd : 32;              // lanes along the head dimension
    const int nsp = 32 / ld;                       // slot partitions
    const int sp = lane / ld, d0 = lane - sp * ld;
    for (int d = d0; d < hd; d += ld) {
        float acc = 0.f;
        for (int slot = sp; slot < L; slot += nsp) acc = fmaf(pbuf[slot], Vr[(size_t)slot * E + d], acc);
        for (int o = ld; o < 32; o <<= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (sp == 0) out[d * os] = acc;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// K1 (generic): one CTA per group of G trajectories, whole horizon in one launch.
// Activations live in shared memory, transposed [channel][g]; weights stream from L2; the KV window of
// the group is a ring in the global workspace.  Handles any n_embd (multiple of 32) / head size.
// ------------------------------------------------------------------------------------------------
struct RolloutParams {
    const float* blob;
    const float* blob_sw;    // XOR-swizzled weight image for the tensor-core kernel (E = 32 only)
    const float* pe;
    const float* x0;
    long long x0_stride;
    float* out;
    long long out_stride;
    float* kv;
    float* presents;
    int B, S, use_cache;
    float eps;
    Gpt2Layout lay;
    int use_tma;
    int prefetch;
    int spin_ns;             // back-off between mbarrier probes of the warp-specialised kernel
    long long* prof;         // WS_PROF builds only: per-warp (wait, total) cycles of CTA 0
};

template <int G>
__device__ __forceinline__ void block_ln_T(const float* xT, float* outT, const float* __restrict__ w,
                                           const float* __restrict__ b, int E, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int g = warp; g < G; g += nw) {
        float s = 0.f;
        for (int c = lane; c < E; c += 32) s += xT[c * G + g];
        const float mean = warp_sum(s) / (float)E;
        float v = 0.f;
        for (int c = lane; c < E; c += 32) {
            const float d = xT[c * G + g] - mean;
            v = fmaf(d, d, v);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)E + eps);
        for (int c = lane; c < E; c += 32)
            outT[c * G + g] = (xT[c * G + g] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
    }
}

template <int G>
__global__ void __launch_bounds__(ROLL_THREADS) rollout_generic_kernel(RolloutParams p) {
    extern __shared__ __align__(16) float sm[];
    const Gpt2Layout& L = p.lay;
    const int E = L.E, H = L.n_head, hd = L.hd, C = L.n_ctx;
    float* xT = sm;                       // [E][G]   residual stream
    float* aT = xT + E * G;               // [E][G]   gemv input / attention output
    float* qT = aT + E * G;               // [3E][G]  q | k | v
    float* hT = qT + 3 * E * G;           // [4E][G]  MLP hidden
    float* part = hT + 4 * E * G;         // [4*blockDim*G]
    float* pbuf = part + 4 * ROLL_THREADS * G; // [nwarps][C]
    const int tid = threadIdx.x, warp = tid >> 5, nw = blockDim.x >> 5;
    const int b0 = blockIdx.x * G;
    const size_t kv_per_g = (size_t)L.n_layer * 2 * C * E;
    float* kvb = p.kv + (size_t)blockIdx.x * G * kv_per_g;

    for (int i = tid; i < E * G; i += blockDim.x) {
        const int c = i / G, g = i - c * G;
        const int b = b0 + g;
        const float x = (b < p.B) ? p.x0[(size_t)b * p.x0_stride + c] : 0.f;
        xT[i] = x + __ldg(p.pe + c);      // position 0
    }
    __syncthreads();

    for (int s = 0; s < p.S; s++) {
        const int Lk = p.use_cache ? min(s + 1, C) : 1;
        const int slot = p.use_cache ? (s % C) : 0;
        for (int l = 0; l < L.n_layer; l++) {
            const float* w = p.blob + (size_t)l * L.layer_stride;
            block_ln_T<G>(xT, aT, w + L.ln1w, w + L.ln1b, E, p.eps);
            __syncthreads();
            block_gemv4<G>(w + L.wqkv, w + L.bqkv, E, 3 * E, aT, part,
                          [&](int n, int g, float v) { qT[n * G + g] = v; });
            __syncthreads();
            // append (k, v) to the ring
            for (int i = tid; i < 2 * E * G; i += blockDim.x) {
                const int g = i / (2 * E), r = i - g * 2 * E;
                const int which = r / E, e = r - which * E;
                kvb[(size_t)g * kv_per_g + ((size_t)(l * 2 + which) * C + slot) * E + e] =
                    qT[((1 + which) * E + e) * G + g];
            }
            __syncthreads();
            for (int pair = warp; pair < G * H; pair += nw) {
                const int g = pair / H, hh = pair - g * H;
                const float* Kr = kvb + (size_t)g * kv_per_g + (size_t)(l * 2 + 0) * C * E + hh * hd;
                const float* Vr = kvb + (size_t)g * kv_per_g + (size_t)(l * 2 + 1) * C * E + hh * hd;
                warp_attention(qT + (hh * hd) * G + g, G, Kr, Vr, E, hd, Lk, pbuf + warp * C,
                               aT + (hh * hd) * G + g, G, nullptr);
            }
            __syncthreads();
            block_gemv4<G>(w + L.wproj, w + L.bproj, E, E, aT, part,
                          [&](int n, int g, float v) { xT[n * G + g] += v; });
            __syncthreads();
            block_ln_T<G>(xT, aT, w + L.ln2w, w + L.ln2b, E, p.eps);
            __syncthreads();
            block_gemv4<G>(w + L.wfc, w + L.bfc, E, 4 * E, aT, part,
                          [&](int n, int g, float v) { hT[n * G + g] = gelu_new(v); });
            __syncthreads();
            block_gemv4<G>(w + L.wpr, w + L.bpr, 4 * E, E, hT, part,
                          [&](int n, int g, float v) { xT[n * G + g] += v; });
            __syncthreads();
        }
        block_ln_T<G>(xT, aT, p.blob + L.lnfw, p.blob + L.lnfb, E, p.eps);
        __syncthreads();
        const int pnext = p.use_cache ? min(s + 1, C - 1) : 0;
        block_gemv4<G>(p.blob + L.wf, p.blob + L.bf, E, E, aT, part, [&](int n, int g, float v) {
            const int b = b0 + g;
            if (b < p.B) p.out[(size_t)b * p.out_stride + (size_t)s * E + n] = v;
            xT[n * G + g] = v + __ldg(p.pe + pnext * E + n);
        });
        __syncthreads();
    }

    if (p.presents != nullptr && p.use_cache) {
        // presents[l][which][b][h][t][d], chronological t; entry s lives in ring slot s % C
        const int Tk = min(p.S, C);
        const size_t per_l = (size_t)2 * p.B * H * Tk * hd;
        for (int g = 0; g < G; g++) {
            const int b = b0 + g;
            if (b >= p.B) break;
            for (int i = tid; i < L.n_layer * 2 * Tk * E; i += blockDim.x) {
                const int e = i % E;
                int r = i / E;
                const int t = r % Tk;
                r /= Tk;
                const int which = r & 1, l = r >> 1;
                const int sl = (p.S - Tk + t) % C;
                const int hh = e / hd, d = e - hh * hd;
                p.presents[(size_t)l * per_l + ((((size_t)which * p.B + b) * H + hh) * Tk + t) * hd + d] =
                    kvb[(size_t)g * kv_per_g + ((size_t)(l * 2 + which) * C + sl) * E + e];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1 (E = 32, head size 8): persistent warp-autonomous rollout.
//   * every weight of the model (<= 227 KB) is staged ONCE per CTA into shared memory by TMA bulk copies;
//   * one warp owns G trajectories for the whole horizon: lane n owns channel n of the residual stream
//     (registers), GEMVs are register-blocked over the G trajectories with weights read from smem
//     (one conflict-free row per k) and activations broadcast from a per-warp transposed tile;
//   * no CTA-wide barrier inside the step loop (only __syncwarp);
//   * the KV window of the G trajectories is a ring in a per-warp slice of the global workspace that is
//     reused wave after wave, so it stays L2-resident; window reads are 256-bit loads.
// ------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void load_a(const float* s, float (&a)[G]) {
    if constexpr (G == 8) {
        const float4 u = *reinterpret_cast<const float4*>(s);
        const float4 v = *reinterpret_cast<const float4*>(s + 4);
        a[0] = u.x; a[1] = u.y; a[2] = u.z; a[3] = u.w; a[4] = v.x; a[5] = v.y; a[6] = v.z; a[7] = v.w;
    } else if constexpr (G == 4) {
        const float4 u = *reinterpret_cast<const float4*>(s);
        a[0] = u.x; a[1] = u.y; a[2] = u.z; a[3] = u.w;
    } else if constexpr (G == 2) {
        const float2 u = *reinterpret_cast<const float2*>(s);
        a[0] = u.x; a[1] = u.y;
    } else {
        a[0] = s[0];
    }
}

template <int G>
__device__ __forceinline__ void store_T(float* scr, int row, const float (&x)[G]) {
    float* d = scr + row * G;
    if constexpr (G == 8) {
        *reinterpret_cast<float4*>(d) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(x[4], x[5], x[6], x[7]);
    } else if constexpr (G == 4) {
        *reinterpret_cast<float4*>(d) = make_float4(x[0], x[1], x[2], x[3]);
    } else if constexpr (G == 2) {
        *reinterpret_cast<float2*>(d) = make_float2(x[0], x[1]);
    } else {
        d[0] = x[0];
    }
}

// LayerNorm over the 32 channels of each of the G trajectories, in place on the transposed tile
// scr[c*G+g].  Lane `lane` touches flat elements i*32+lane (conflict-free); its trajectory is lane % G.
template <int G>
__device__ __forceinline__ void warp_ln_T(float* scr, const float* lw, const float* lb, float eps, int lane) {
    float x[G];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < G; i++) {
        x[i] = scr[i * 32 + lane];
        s += x[i];
    }
#pragma unroll
    for (int o = G; o < 32; o <<= 1) s += __shfl_xor_sync(FULL, s, o);
    const float mean = s * (1.0f / 32.0f);
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < G; i++) {
        x[i] -= mean;
        v = fmaf(x[i], x[i], v);
    }
#pragma unroll
    for (int o = G; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
    const float rstd = 1.0f / sqrtf(v * (1.0f / 32.0f) + eps);
#pragma unroll
    for (int i = 0; i < G; i++) {
        const int c = (i * 32 + lane) / G;
        scr[i * 32 + lane] = x[i] * rstd * lw[c] + lb[c];
    }
}

// acc[c][g] += sum_k scr[k*G+g] * W[k*ldw + lane + 32*c]
template <int G, int C>
__device__ __forceinline__ void warp_gemv(const float* scr, const float* W, int K, int ldw, int lane,
                                          float (&acc)[C][G]) {
#pragma unroll 8                 // measured: 8 beats 4 by 4-22 % (loads of 8 k in flight per lane), 16 and a [K/4][N][4] packing with 128-bit weight loads lose
    for (int k = 0; k < K; k++) {
        float a[G];
        load_a<G>(scr + k * G, a);
        float w[C];
#pragma unroll
        for (int c = 0; c < C; c++) w[c] = W[k * ldw + lane + 32 * c];
#pragma unroll
        for (int c = 0; c < C; c++) {
            if constexpr (G % 2 == 0) {
#pragma unroll
                for (int g = 0; g < G; g += 2) ffma2(acc[c][g], acc[c][g + 1], a[g], a[g + 1], w[c]);
            } else {
#pragma unroll
                for (int g = 0; g < G; g++) acc[c][g] = fmaf(a[g], w[c], acc[c][g]);
            }
        }
    }
}

// Attention for the G trajectories of a warp, 4 heads of 8 (hd = 8), window of L <= NCTX keys.
// lane = ((g*4 + head) * P + part) with P = 8/G: for G = 8 every lane owns one (trajectory, head) pair and walks
// the whole window alone — no shuffles at all; for smaller G the P lanes of a pair split the slots and combine
// max / sum / output with P-lane xor-shuffles.  All key (then value) rows of a batch are loaded before they are
// consumed (independent 256-bit loads: memory-level parallelism instead of a load->use chain per slot).
//   qs   : smem [G][32] queries (row per trajectory)
//   kvl  : ring of this warp for this layer: trajectory g at kvl + g*kv_per_g, K rows [C][32] then V rows [C][32]
//   outT : smem transposed tile [32][G] receiving the merged-head output (input of the projection GEMV)
// rows base, base + P*128 B, ... with the offsets folded into the load instructions
template <int BS, int P, int JJ = 0>
__device__ __forceinline__ void load_rows(const float* base, float (&r)[BS][8]) {
    if constexpr (JJ < BS) {
        ldg256_off<JJ * P * 128>(base, r[JJ]);
        load_rows<BS, P, JJ + 1>(base, r);
    }
}

// FULLW: the window is full (L == C == NCTX, every step from n_ctx-1 on): no row clamping, no validity selects.
// QS: row stride of the query tile; ORM > 0: write the output row-major [g][32] with row stride ORM instead of the
// transposed tile.
template <int G, int NCTX, bool FULLW, int QS = 32, int ORM = 0>
__device__ __forceinline__ void warp_attn_e32(const float* qs, const float* kvl, size_t kv_per_g, int C, int L,
                                              float* outT, int lane) {
    constexpr int P = 8 / G;
    constexpr int NS = NCTX / P;                      // slots per lane
    constexpr int BS = (G >= 4 && NS <= 32) ? (NS < 16 ? NS : 16) : (NS >= 64 ? 4 : (NS < 8 ? NS : 8));   // rows in flight per lane
    const int part = lane % P, gh = lane / P, hh = gh & 3, g = gh >> 2;
    float q[8];
    {
        const float4* qp = reinterpret_cast<const float4*>(qs + g * QS + hh * 8);
        const float4 a = qp[0], b = qp[1];
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    __syncwarp();                                     // every lane holds its query: the tile may be overwritten
    const float* Kp = kvl + (size_t)g * kv_per_g + hh * 8;
    const float* Vp = Kp + C * 32;
    // Scores are kept in the log2 domain: the query is pre-multiplied ONCE by log2(e)/sqrt(8) (attention.py:99 divides
    // q.k by sqrt(head size)), so a window row costs its 8 FMAs, one max, and later one subtract + ex2.approx; the
    // softmax denominator is applied to the 8 outputs instead of the 32 weights.  Each step is <= 2 ulp from the
    // reference's divide / expf / normalise-then-weight form (parity tests: unchanged at ~5e-7 against the 1e-5 bar).
#pragma unroll
    for (int d = 0; d < 8; d++) q[d] *= 0.35355339059327373f * 1.4426950408889634f;
    float sc[NS];
    float m = -INFINITY;
#pragma unroll
    for (int j0 = 0; j0 < NS; j0 += BS) {
        if (FULLW || j0 * P < L) {                    // warp-uniform
            float kk[BS][8];
            if constexpr (FULLW) {
                load_rows<BS, P>(Kp + (j0 * P + part) * 32, kk);
            } else {
#pragma unroll
                for (int jj = 0; jj < BS; jj++) {
                    const int slot = (j0 + jj) * P + part;
                    ldg256(Kp + min(slot, L - 1) * 32, kk[jj]);
                }
            }
#pragma unroll
            for (int jj = 0; jj < BS; jj++) {
                const int slot = (j0 + jj) * P + part;
                float s0 = 0.f;
#pragma unroll
                for (int d = 0; d < 8; d++) s0 = fmaf(q[d], kk[jj][d], s0);
                const float s1 = (FULLW || slot < L) ? s0 : -INFINITY;
                sc[j0 + jj] = s1;
                m = fmaxf(m, s1);
            }
        } else {
#pragma unroll
            for (int jj = 0; jj < BS; jj++) sc[j0 + jj] = -INFINITY;
        }
    }
#pragma unroll
    for (int o = 1; o < P; o <<= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const float pj = (FULLW || j * P + part < L) ? ex2_approx(sc[j] - m) : 0.f;
        sc[j] = pj;
        sum += pj;
    }
#pragma unroll
    for (int o = 1; o < P; o <<= 1) sum += __shfl_xor_sync(FULL, sum, o);
    const float rs = 1.0f / sum;
    float acc[8];
#pragma unroll
    for (int d = 0; d < 8; d++) acc[d] = 0.f;
#pragma unroll
    for (int j0 = 0; j0 < NS; j0 += BS) {
        if (FULLW || j0 * P < L) {
            float vv[BS][8];
            if constexpr (FULLW) {
                load_rows<BS, P>(Vp + (j0 * P + part) * 32, vv);
            } else {
#pragma unroll
                for (int jj = 0; jj < BS; jj++) {
                    const int slot = (j0 + jj) * P + part;
                    ldg256(Vp + min(slot, L - 1) * 32, vv[jj]);       // clamped rows carry weight 0
                }
            }
#pragma unroll
            for (int jj = 0; jj < BS; jj++) {
                const float pw = sc[j0 + jj];
#pragma unroll
                for (int d = 0; d < 8; d += 2) ffma2(acc[d], acc[d + 1], vv[jj][d], vv[jj][d + 1], pw);
            }
        }
    }
#pragma unroll
    for (int o = 1; o < P; o <<= 1)
#pragma unroll
        for (int d = 0; d < 8; d++) acc[d] += __shfl_xor_sync(FULL, acc[d], o);
#pragma unroll
    for (int d = 0; d < 8; d++) acc[d] *= rs;
    if (part == 0) {
        if constexpr (ORM > 0) {
            *reinterpret_cast<float4*>(outT + g * ORM + hh * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(outT + g * ORM + hh * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        } else {
#pragma unroll
            for (int d = 0; d < 8; d++) outT[(hh * 8 + d) * G + g] = acc[d];
        }
    }
}

template <int G, int MAXI, bool CACHE>
#ifndef TRPX_E32_CACHE_NT
#define TRPX_E32_CACHE_NT 256
#endif
__global__ void __launch_bounds__((CACHE && G >= 4) ? TRPX_E32_CACHE_NT : 384, 1) rollout_e32_kernel(RolloutParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t bar;
    const Gpt2Layout& L = p.lay;
    const int C = L.n_ctx, NL = L.n_layer;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, W = blockDim.x >> 5;
    float* wsm = sm;
    float* scr = sm + L.total + warp * (64 * G);

    // ---- stage all weights into shared memory ----
    if (p.use_tma) {
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)L.total * 4u;
            mbar_expect_tx(&bar, bytes);
            const uint32_t CH = 32768u;
            for (uint32_t off = 0; off < bytes; off += CH) {
                const uint32_t n = min(CH, bytes - off);
                bulk_g2s(reinterpret_cast<char*>(wsm) + off, reinterpret_cast<const char*>(p.blob) + off, n, &bar);
            }
        }
        mbar_wait(&bar, 0);
    } else {
        const float4* src = reinterpret_cast<const float4*>(p.blob);
        float4* dst = reinterpret_cast<float4*>(wsm);
        for (int i = tid; i < L.total / 4; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
    }

    const int ngroups = (p.B + G - 1) / G;
    const size_t kv_per_g = (size_t)NL * 2 * C * 32;
    float* kvw = CACHE ? p.kv + (size_t)(blockIdx.x * W + warp) * G * kv_per_g : nullptr;
    const float* gw = wsm + L.lnfw;     // globals: lnfw, lnfb, wf, bf are consecutive

    for (int grp = blockIdx.x * W + warp; grp < ngroups; grp += gridDim.x * W) {
        const int b0 = grp * G;
        float h[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int b = b0 + g;
            h[g] = (b < p.B ? p.x0[(size_t)b * p.x0_stride + lane] : 0.f) + __ldg(p.pe + lane);
        }
        for (int s = 0; s < p.S; s++) {
            const int Lk = CACHE ? min(s + 1, C) : 1;
            const int slot = CACHE ? (s % C) : 0;
            for (int l = 0; l < NL; l++) {
                const float* w = wsm + l * L.layer_stride;
                if constexpr (CACHE) {
                    // Pull this layer's window into L2 now: one bulk prefetch per trajectory for K and for V (no
                    // registers, no wait).  The HBM latency then overlaps LN1 + the QKV GEMV instead of being
                    // exposed inside the attention; the prefetched data only has to survive that short phase
                    // (prefetching a whole layer ahead was measured to thrash L2 and was slower).
                    if (p.prefetch == 1 && lane < G && Lk > 1) {
                        const float* kp = kvw + (size_t)lane * kv_per_g + (size_t)(l * 2) * C * 32;
                        prefetch_l2_bulk(kp, (uint32_t)Lk * 128u);
                        prefetch_l2_bulk(kp + C * 32, (uint32_t)Lk * 128u);
                    }
                }
                // ---- LN1 ----
                store_T<G>(scr, lane, h);
                __syncwarp();
                warp_ln_T<G>(scr, w + L.ln1w, w + L.ln1b, p.eps, lane);
                __syncwarp();
                if constexpr (CACHE) {
                    float qkv[3][G];
#pragma unroll
                    for (int c = 0; c < 3; c++)
#pragma unroll
                        for (int g = 0; g < G; g++) qkv[c][g] = w[L.bqkv + lane + 32 * c];
                    warp_gemv<G, 3>(scr, w + L.wqkv, 32, 96, lane, qkv);
                    // ---- append (k, v) to the ring, then attend over the window ----
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        float* kp = kvw + (size_t)g * kv_per_g + (size_t)(l * 2) * C * 32;
                        kp[slot * 32 + lane] = qkv[1][g];
                        kp[C * 32 + slot * 32 + lane] = qkv[2][g];
                    }
                    __syncwarp();                 // LN tile fully consumed; ring stores visible to the warp
#pragma unroll
                    for (int g = 0; g < G; g++) scr[g * 32 + lane] = qkv[0][g];
                    __syncwarp();
                    if (Lk == 8 * MAXI)
                        warp_attn_e32<G, 8 * MAXI, true>(scr, kvw + (size_t)(l * 2) * C * 32, kv_per_g, C, Lk, scr, lane);
                    else
                        warp_attn_e32<G, 8 * MAXI, false>(scr, kvw + (size_t)(l * 2) * C * 32, kv_per_g, C, Lk, scr, lane);
                    if (p.prefetch == 2 && lane < G) {
                        // mid-layer distance: the NEXT attention's window (next layer, or layer 0 of the next step)
                        // is requested now and has the projection + MLP phase to arrive
                        const bool wrap = (l + 1 == NL);
                        const int ln = wrap ? 0 : l + 1;
                        const int Ln = wrap ? min(s + 2, C) : Lk;
                        if ((!wrap || s + 1 < p.S) && Ln > 1) {
                            const float* kp = kvw + (size_t)lane * kv_per_g + (size_t)(ln * 2) * C * 32;
                            prefetch_l2_bulk(kp, (uint32_t)Ln * 128u);
                            prefetch_l2_bulk(kp + C * 32, (uint32_t)Ln * 128u);
                        }
                    }
                    __syncwarp();
                } else {
                    // window of one key: softmax == 1 exactly, so the head output is v itself
                    float vv[1][G];
#pragma unroll
                    for (int g = 0; g < G; g++) vv[0][g] = w[L.bqkv + 64 + lane];
                    warp_gemv<G, 1>(scr, w + L.wqkv + 64, 32, 96, lane, vv);
                    __syncwarp();
                    store_T<G>(scr, lane, vv[0]);
                    __syncwarp();
                }
                // ---- output projection + residual ----
                {
                    float acc[1][G];
#pragma unroll
                    for (int g = 0; g < G; g++) acc[0][g] = w[L.bproj + lane];
                    warp_gemv<G, 1>(scr, w + L.wproj, 32, 32, lane, acc);
#pragma unroll
                    for (int g = 0; g < G; g++) h[g] += acc[0][g];
                }
                // ---- LN2 + MLP ----
                __syncwarp();
                store_T<G>(scr, lane, h);
                __syncwarp();
                warp_ln_T<G>(scr, w + L.ln2w, w + L.ln2b, p.eps, lane);
                __syncwarp();
                float f[4][G];
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int g = 0; g < G; g++) f[c][g] = w[L.bfc + lane + 32 * c];
                warp_gemv<G, 4>(scr, w + L.wfc, 32, 128, lane, f);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if constexpr (G % 2 == 0) {
#pragma unroll
                        for (int g = 0; g < G; g += 2) gelu_new_sig2(f[c][g], f[c][g + 1]);
                    } else {
                        f[c][0] = gelu_new_sig(f[c][0]);
                    }
                }
                float acc[1][G];
#pragma unroll
                for (int g = 0; g < G; g++) acc[0][g] = w[L.bpr + lane];
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    __syncwarp();
                    store_T<G>(scr, lane, f[2 * half]);
                    store_T<G>(scr, lane + 32, f[2 * half + 1]);
                    __syncwarp();
                    warp_gemv<G, 1>(scr, w + L.wpr + half * 64 * 32, 64, 32, lane, acc);
                }
#pragma unroll
                for (int g = 0; g < G; g++) h[g] += acc[0][g];
                __syncwarp();
            }
            // ---- final LN + linear head ----
            store_T<G>(scr, lane, h);
            __syncwarp();
            warp_ln_T<G>(scr, gw, gw + 32, p.eps, lane);
            __syncwarp();
            float y[1][G];
#pragma unroll
            for (int g = 0; g < G; g++) y[0][g] = gw[64 + 1024 + lane];
            warp_gemv<G, 1>(scr, gw + 64, 32, 32, lane, y);
            const int pnext = CACHE ? min(s + 1, C - 1) : 0;
            const float pe = __ldg(p.pe + pnext * 32 + lane);
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int b = b0 + g;
                if (b < p.B) p.out[(size_t)b * p.out_stride + (size_t)s * 32 + lane] = y[0][g];
                h[g] = y[0][g] + pe;
            }
            __syncwarp();
        }
        if (CACHE && p.presents != nullptr) {
            const int Tk = min(p.S, C);
            const int H = 4, hd = 8;
            const size_t per_l = (size_t)2 * p.B * H * Tk * hd;
            const int hh = lane >> 3, d = lane & 7;
            for (int g = 0; g < G; g++) {
                const int b = b0 + g;
                if (b >= p.B) break;
                for (int r = 0; r < NL * 2 * Tk; r++) {
                    const int t = r % Tk, lw = r / Tk;          // lw = l*2 + which
                    const int which = lw & 1, l = lw >> 1;
                    const int sl = (p.S - Tk + t) % C;
                    p.presents[(size_t)l * per_l + ((((size_t)which * p.B + b) * H + hh) * Tk + t) * hd + d] =
                        kvw[(size_t)g * kv_per_g + ((size_t)lw * C + sl) * 32 + lane];
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1 (E = 32, KV window), TMA-staged attention + per-layer weight ring.
// Measured on rollout_e32_kernel<4,...> (profiles/README.md): the bare window stream runs at the HBM peak
// (trpx_probe_window_stream: 6.4 TB/s with the same 8 warps x 16 rows in flight) and the dense part alone takes
// 56 ms, yet the kernel takes their SUM (130 ms): every warp is a latency chain that holds 128 registers of
// in-flight window rows, so only 8 warps fit and nothing hides a warp's wait.  Here the window never touches the
// LSU or the register file:
//   * a warp owns 4 trajectories and a 16 KB shared-memory stage [4][BR = 32 rows][32 floats];
//   * one elected lane issues 4 TMA bulk copies (cp.async.bulk, one contiguous window slice per trajectory) that
//     complete on the warp's own mbarrier; the key stage of the NEXT layer is issued as soon as this layer's
//     attention is done, so it lands during projection + MLP + LN + QKV; value rows (and key rows beyond the first
//     stage, n_ctx = 64) are bulk-prefetched into L2 at the same moment and staged when the buffer frees up;
//   * the row of the current token is patched into the stage from registers (its global copy may be stale);
//   * lanes read the stage with LDS.128, the two lanes of a (trajectory, head) pair in opposite half order so
//     that a quarter-warp phase never hits a bank twice;
//   * to make room for the stages the weights are no longer resident as a whole (203 KB): a producer warp streams
//     one layer (50.8 KB) at a time from L2 into a two-slot ring (full/empty mbarriers, TMA bulk copies), the
//     consumer warps run at most one layer apart; the head (ln_f, mlp_f) stays resident;
//   * arithmetic order is the one of rollout_e32_kernel (bit-identical results).
// ------------------------------------------------------------------------------------------------
constexpr int E32S_G = 4;
constexpr int E32S_MAXW = 12;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NCTX>
__global__ void __launch_bounds__(NCTX >= 32 ? 256 : 416, 1) rollout_e32s_kernel(RolloutParams p) {
    constexpr int G = E32S_G, P = 2;
    constexpr int BR = NCTX < 32 ? NCTX : 32;         // rows per stage
    constexpr int NS = NCTX / P;                      // scores per lane
    constexpr int JB = BR / P;                        // rows per lane per stage
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t hbar;            // head weights
    __shared__ __align__(8) uint64_t full[2], empty[2];
    __shared__ __align__(8) uint64_t tbar_all[E32S_MAXW];
    const Gpt2Layout& L = p.lay;
    const int C = L.n_ctx, NL = L.n_layer;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = (blockDim.x >> 5) - 1;              // consumer warps; warp W is the weight producer
    const int LS = L.layer_stride;
    const int head_floats = L.total - L.lnfw;         // ln_f w, b, mlp_f W^T, b: consecutive at the end of the blob
    float* gw = sm;                                   // resident head
    float* wring = sm + ((head_floats + 3) & ~3);     // [2][LS]
    float* scr = wring + 2 * LS + warp * (64 * G + G * BR * 32);
    float* buf = scr + 64 * G;
    uint64_t* tbar = &tbar_all[warp < E32S_MAXW ? warp : 0];

    const int ngroups = (p.B + G - 1) / G;
    const int iters = (ngroups + gridDim.x * W - 1) / (gridDim.x * W);    // same for every warp of the CTA
    const int total_layers = iters * p.S * NL;

    if (tid == 0) {
        mbar_init(&hbar, 1);
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&empty[0], W);
        mbar_init(&empty[1], W);
    }
    if (lane == 0 && warp < W) mbar_init(tbar, 1);
    __syncthreads();
    if (warp == W) {
        // ---------------- producer: head once, then layer after layer into the two-slot ring ----------------
        if (lane == 0) {
            mbar_expect_tx(&hbar, (uint32_t)head_floats * 4u);
            bulk_g2s(gw, p.blob + L.lnfw, (uint32_t)head_floats * 4u, &hbar);
            const uint32_t bytes = (uint32_t)LS * 4u;
            for (int lc = 0; lc < total_layers; lc++) {
                const int b = lc & 1, u = lc >> 1;
                if (u > 0) mbar_wait_backoff(&empty[b], (u - 1) & 1, 256);
                mbar_expect_tx(&full[b], bytes);
                const char* src = reinterpret_cast<const char*>(p.blob + (size_t)(lc % NL) * LS);
                char* dst = reinterpret_cast<char*>(wring + b * LS);
                const uint32_t CH = 16384u;
                for (uint32_t off = 0; off < bytes; off += CH) bulk_g2s(dst + off, src + off, min(CH, bytes - off), &full[b]);
            }
        }
        return;
    }
    mbar_wait(&hbar, 0);

    const size_t kv_per_g = (size_t)NL * 2 * C * 32;
    float* kvw = p.kv + (size_t)(blockIdx.x * W + warp) * G * kv_per_g;
    const int part = lane & 1, gh = lane >> 1, hh = gh & 3, ga = gh >> 2;      // attention role of the lane
    uint32_t par = 0;                                 // parity of the next completion of tbar
    int lc = 0;                                       // layers consumed so far (ring position)

    // stage rows [stg*BR, min(Lk, (stg+1)*BR)) of the key (which = 0) / value (1) windows of layer l
    auto issue = [&](int l, int which, int stg, int Lk) {
        if (lane == 0) {
            const int r0 = stg * BR;
            const uint32_t nb = (uint32_t)min(Lk - r0, BR) * 128u;
            mbar_expect_tx(tbar, nb * G);
#pragma unroll
            for (int g = 0; g < G; g++)
                bulk_g2s(buf + g * BR * 32, kvw + (size_t)g * kv_per_g + ((size_t)(l * 2 + which) * C + r0) * 32, nb, tbar);
        }
    };
    // everything of layer l's window that is not in the first key stage goes to L2 now
    auto prefetch_rest = [&](int l, int Lk) {
        if (p.prefetch == 1 && lane < G) {
            const float* kp = kvw + (size_t)lane * kv_per_g + (size_t)(l * 2) * C * 32;
            if (Lk > BR) prefetch_l2_bulk(kp + BR * 32, (uint32_t)(Lk - BR) * 128u);
            prefetch_l2_bulk(kp + C * 32, (uint32_t)Lk * 128u);
        }
    };
    // the 8 floats of (trajectory ga, head hh) in stage row r; the two lanes of a pair read the halves in opposite order
    auto lds_row = [&](int r, float (&x)[8]) {
        const float4* rp = reinterpret_cast<const float4*>(buf + (ga * BR + r) * 32 + hh * 8);
        const float4 a = rp[part], b = rp[part ^ 1];
        const float4 lo = part ? b : a, hi = part ? a : b;
        x[0] = lo.x; x[1] = lo.y; x[2] = lo.z; x[3] = lo.w; x[4] = hi.x; x[5] = hi.y; x[6] = hi.z; x[7] = hi.w;
    };

    for (int it = 0; it < iters; it++) {
        const int grp = (it * gridDim.x + blockIdx.x) * W + warp;
        if (grp >= ngroups) {
            // ragged tail: keep the ring protocol going without computing
            for (int i = 0; i < p.S * NL; i++, lc++) {
                mbar_wait(&full[lc & 1], (lc >> 1) & 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[lc & 1]);
            }
            continue;
        }
        const int b0 = grp * G;
        float h[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int b = b0 + g;
            h[g] = (b < p.B ? p.x0[(size_t)b * p.x0_stride + lane] : 0.f) + __ldg(p.pe + lane);
        }
        bool staged = false;                          // first key stage of the upcoming attention already in flight?
        for (int s = 0; s < p.S; s++) {
            const int Lk = min(s + 1, C);
            const int slot = s % C;
            for (int l = 0; l < NL; l++, lc++) {
                if (!staged && Lk > 1) {
                    issue(l, 0, 0, Lk);
                    prefetch_rest(l, Lk);
                    staged = true;
                }
                mbar_wait(&full[lc & 1], (lc >> 1) & 1);
                const float* w = wring + (lc & 1) * LS;
                // ---- LN1 + QKV ----
                store_T<G>(scr, lane, h);
                __syncwarp();
                warp_ln_T<G>(scr, w + L.ln1w, w + L.ln1b, p.eps, lane);
                __syncwarp();
                float qkv[3][G];
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int g = 0; g < G; g++) qkv[c][g] = w[L.bqkv + lane + 32 * c];
                warp_gemv<G, 3>(scr, w + L.wqkv, 32, 96, lane, qkv);
                // ---- append (k, v) to the ring (for the following steps) ----
#pragma unroll
                for (int g = 0; g < G; g++) {
                    float* kp = kvw + (size_t)g * kv_per_g + (size_t)(l * 2) * C * 32;
                    kp[slot * 32 + lane] = qkv[1][g];
                    kp[C * 32 + slot * 32 + lane] = qkv[2][g];
                }
                asm volatile("fence.proxy.async.global;" ::: "memory");   // later TMA reads of these rows
                __syncwarp();                         // LN tile fully consumed
#pragma unroll
                for (int g = 0; g < G; g++) scr[g * 32 + lane] = qkv[0][g];
                __syncwarp();
                float q[8];
                {
                    const float4* qp = reinterpret_cast<const float4*>(scr + ga * 32 + hh * 8);
                    const float4 a = qp[0], b = qp[1];
                    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
                }
#pragma unroll
                for (int d = 0; d < 8; d++) q[d] *= 0.35355339059327373f * 1.4426950408889634f;   // as warp_attn_e32
                // ---- scores (log2 domain), stage by stage ----
                float sc[NS];
                float m = -INFINITY;
#pragma unroll
                for (int stg = 0; stg < NCTX / BR; stg++) {
                    if (stg * BR < Lk) {              // warp-uniform
                        if (Lk > 1) {
                            if (stg > 0) {
                                __syncwarp();
                                issue(l, 0, stg, Lk);
                            }
                            mbar_wait(tbar, par);
                            par ^= 1;
                        }
                        const int ps = slot - stg * BR;
                        if (ps >= 0 && ps < BR) {
#pragma unroll
                            for (int g = 0; g < G; g++) buf[(g * BR + ps) * 32 + lane] = qkv[1][g];
                        }
                        __syncwarp();
#pragma unroll
                        for (int jj = 0; jj < JB; jj++) {
                            const int row = jj * P + part;
                            float kk[8];
                            lds_row(min(row, min(Lk - stg * BR, BR) - 1), kk);
                            float s0 = 0.f;
#pragma unroll
                            for (int d = 0; d < 8; d++) s0 = fmaf(q[d], kk[d], s0);
                            const float s1 = (stg * BR + row < Lk) ? s0 : -INFINITY;
                            sc[stg * JB + jj] = s1;
                            m = fmaxf(m, s1);
                        }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < JB; jj++) sc[stg * JB + jj] = -INFINITY;
                    }
                }
                m = fmaxf(m, __shfl_xor_sync(FULL, m, 1));
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < NS; j++) {
                    const int stg = j / JB, jj = j - stg * JB;
                    const float pj = (stg * BR + jj * P + part < Lk) ? ex2_approx(sc[j] - m) : 0.f;
                    sc[j] = pj;
                    sum += pj;
                }
                sum += __shfl_xor_sync(FULL, sum, 1);
                const float rs = 1.0f / sum;
                float acc[8];
#pragma unroll
                for (int d = 0; d < 8; d++) acc[d] = 0.f;
#pragma unroll
                for (int stg = 0; stg < NCTX / BR; stg++) {
                    if (stg * BR < Lk) {
                        __syncwarp();                 // the stage is fully consumed
                        if (Lk > 1) {
                            issue(l, 1, stg, Lk);
                            mbar_wait(tbar, par);
                            par ^= 1;
                        }
                        const int ps = slot - stg * BR;
                        if (ps >= 0 && ps < BR) {
#pragma unroll
                            for (int g = 0; g < G; g++) buf[(g * BR + ps) * 32 + lane] = qkv[2][g];
                        }
                        __syncwarp();
#pragma unroll
                        for (int jj = 0; jj < JB; jj++) {
                            const int row = jj * P + part;
                            float vv[8];
                            lds_row(min(row, min(Lk - stg * BR, BR) - 1), vv);      // clamped rows carry weight 0
                            const float pw = sc[stg * JB + jj];
#pragma unroll
                            for (int d = 0; d < 8; d++) acc[d] = fmaf(pw, vv[d], acc[d]);
                        }
                    }
                }
                __syncwarp();                         // stage free again: start the next attention's key stage
                {
                    const bool wrap = (l + 1 == NL);
                    const int ln = wrap ? 0 : l + 1;
                    const int Ln = wrap ? min(s + 2, C) : Lk;
                    staged = false;
                    if ((!wrap || s + 1 < p.S) && Ln > 1) {
                        issue(ln, 0, 0, Ln);
                        prefetch_rest(ln, Ln);
                        staged = true;
                    }
                }
#pragma unroll
                for (int d = 0; d < 8; d++) acc[d] += __shfl_xor_sync(FULL, acc[d], 1);
#pragma unroll
                for (int d = 0; d < 8; d++) acc[d] *= rs;
                if (part == 0) {
#pragma unroll
                    for (int d = 0; d < 8; d++) scr[(hh * 8 + d) * G + ga] = acc[d];
                }
                __syncwarp();
                // ---- output projection + residual ----
                {
                    float pr[1][G];
#pragma unroll
                    for (int g = 0; g < G; g++) pr[0][g] = w[L.bproj + lane];
                    warp_gemv<G, 1>(scr, w + L.wproj, 32, 32, lane, pr);
#pragma unroll
                    for (int g = 0; g < G; g++) h[g] += pr[0][g];
                }
                // ---- LN2 + MLP ----
                __syncwarp();
                store_T<G>(scr, lane, h);
                __syncwarp();
                warp_ln_T<G>(scr, w + L.ln2w, w + L.ln2b, p.eps, lane);
                __syncwarp();
                float f[4][G];
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int g = 0; g < G; g++) f[c][g] = w[L.bfc + lane + 32 * c];
                warp_gemv<G, 4>(scr, w + L.wfc, 32, 128, lane, f);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if constexpr (G % 2 == 0) {
#pragma unroll
                        for (int g = 0; g < G; g += 2) gelu_new_sig2(f[c][g], f[c][g + 1]);
                    } else {
                        f[c][0] = gelu_new_sig(f[c][0]);
                    }
                }
                float mo[1][G];
#pragma unroll
                for (int g = 0; g < G; g++) mo[0][g] = w[L.bpr + lane];
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    __syncwarp();
                    store_T<G>(scr, lane, f[2 * half]);
                    store_T<G>(scr, lane + 32, f[2 * half + 1]);
                    __syncwarp();
                    warp_gemv<G, 1>(scr, w + L.wpr + half * 64 * 32, 64, 32, lane, mo);
                }
#pragma unroll
                for (int g = 0; g < G; g++) h[g] += mo[0][g];
                __syncwarp();                         // this warp is done with the layer's weights
                if (lane == 0) mbar_arrive(&empty[lc & 1]);
            }
            // ---- final LN + linear head ----
            store_T<G>(scr, lane, h);
            __syncwarp();
            warp_ln_T<G>(scr, gw, gw + 32, p.eps, lane);
            __syncwarp();
            float y[1][G];
#pragma unroll
            for (int g = 0; g < G; g++) y[0][g] = gw[64 + 1024 + lane];
            warp_gemv<G, 1>(scr, gw + 64, 32, 32, lane, y);
            const int pnext = min(s + 1, C - 1);
            const float pe = __ldg(p.pe + pnext * 32 + lane);
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int b = b0 + g;
                if (b < p.B) p.out[(size_t)b * p.out_stride + (size_t)s * 32 + lane] = y[0][g];
                h[g] = y[0][g] + pe;
            }
            __syncwarp();
        }
        if (p.presents != nullptr) {
            const int Tk = min(p.S, C);
            const int H = 4, hd = 8;
            const size_t per_l = (size_t)2 * p.B * H * Tk * hd;
            const int ph = lane >> 3, d = lane & 7;
            for (int g = 0; g < G; g++) {
                const int b = b0 + g;
                if (b >= p.B) break;
                for (int r = 0; r < NL * 2 * Tk; r++) {
                    const int t = r % Tk, lw = r / Tk;          // lw = l*2 + which
                    const int which = lw & 1, l = lw >> 1;
                    const int sl = (p.S - Tk + t) % C;
                    p.presents[(size_t)l * per_l + ((((size_t)which * p.B + b) * H + ph) * Tk + t) * hd + d] =
                        kvw[(size_t)g * kv_per_g + ((size_t)lw * C + sl) * 32 + lane];
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1 (E = 32, head size 8), second generation: 2-D lane tiling.
// Measured on the first kernel (ncu, profiles/): its GEMV loop is bound by shared-memory RETURN bandwidth —
// one 32-bit word per lane per cycle per SM — because every lane re-reads the G broadcast activations:
// (G + C) words for G*C FFMAs, i.e. 0.9 (proj, MLP-out) to 2.7 (MLP-in) FFMA per word where >= 4 are needed.
// Here one warp owns 16 trajectories and the 32 lanes form a 4 x 8 grid: lane (r, c) holds the register tile
// [4 trajectories r*4..r*4+3] x [Cc columns], so a k-step costs (4 + Cc) words for 4*Cc FFMAs: 2.0 (Cc = 4),
// 3.0 (QKV, Cc = 12), 3.2 (MLP-in, Cc = 16) — ~1.8x fewer words per FFMA overall.  The residual stream lives in
// the same [4 x 4] tile layout (channels c*4..c*4+3), LayerNorm is done in registers (3 xor-shuffles over c),
// all smem/global accesses of the tile are 128-bit, and each 128-bit weight load of the 8 lanes c covers one
// contiguous 128-byte line (no bank conflicts).  MLP-in column 32*m + 4*c + j sits in lane c's slot (m, j), so
// each quarter m of the hidden units is spread evenly over the lanes for the chunked MLP-out accumulation.
// ------------------------------------------------------------------------------------------------
constexpr int V2G = 16;

// attention of ONE (trajectory, head) pair held entirely by the calling lane; q in registers
template <int NCTX>
__device__ __forceinline__ void lane_attention(const float (&q)[8], const float* Kp, const float* Vp, int L,
                                               float (&acc)[8]) {
    constexpr int BS = NCTX >= 64 ? 4 : 8;
    const float scale = 2.8284271247461903f;
    float sc[NCTX];
    float m = -INFINITY;
#pragma unroll
    for (int j0 = 0; j0 < NCTX; j0 += BS) {
        if (j0 < L) {
            float kk[BS][8];
#pragma unroll
            for (int jj = 0; jj < BS; jj++) ldg256(Kp + min(j0 + jj, L - 1) * 32, kk[jj]);
#pragma unroll
            for (int jj = 0; jj < BS; jj++) {
                float s0 = 0.f;
#pragma unroll
                for (int d = 0; d < 8; d++) s0 = fmaf(q[d], kk[jj][d], s0);
                const float s1 = (j0 + jj < L) ? s0 / scale : -INFINITY;
                sc[j0 + jj] = s1;
                m = fmaxf(m, s1);
            }
        } else {
#pragma unroll
            for (int jj = 0; jj < BS; jj++) sc[j0 + jj] = -INFINITY;
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NCTX; j++) {
        const float pj = (j < L) ? expf(sc[j] - m) : 0.f;
        sc[j] = pj;
        sum += pj;
    }
    const float rs = 1.0f / sum;
#pragma unroll
    for (int d = 0; d < 8; d++) acc[d] = 0.f;
#pragma unroll
    for (int j0 = 0; j0 < NCTX; j0 += BS) {
        if (j0 < L) {
            float vv[BS][8];
#pragma unroll
            for (int jj = 0; jj < BS; jj++) ldg256(Vp + min(j0 + jj, L - 1) * 32, vv[jj]);
#pragma unroll
            for (int jj = 0; jj < BS; jj++) {
                const float pw = sc[j0 + jj] * rs;
#pragma unroll
                for (int d = 0; d < 8; d++) acc[d] = fmaf(pw, vv[jj][d], acc[d]);
            }
        }
    }
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// acc[i][4*m + j] += sum_k aT[k*16 + r*4 + i] * W[k*ldw + MS*m + j]     (NQ float4 weight loads per k; the 8
// lanes c of a row read 8 consecutive float4 = one conflict-free 128-byte line per load)
template <int NQ, int MS = 4>
__device__ __forceinline__ void tile_gemv(const float* aT, const float* W, int K, int ldw, int r,
                                          float (&acc)[4][4 * NQ]) {
#pragma unroll 4
    for (int k = 0; k < K; k++) {
        const float4 a = lds4(aT + k * V2G + r * 4);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int m = 0; m < NQ; m++) {
            const float4 w = lds4(W + k * ldw + MS * m);
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][4 * m + j] = fmaf(av[i], wv[j], acc[i][4 * m + j]);
        }
    }
}

// LayerNorm of the [4 traj][4 ch] register tile over the 32 channels (spread over the 8 lanes c), written
// transposed to aT[(c*4+j)*16 + r*4 + i]
__device__ __forceinline__ void tile_ln_to_smem(const float (&h)[4][4], const float* lw, const float* lb, float eps,
                                                float* aT, int r, int c) {
    const float4 w4 = lds4(lw + c * 4), b4 = lds4(lb + c * 4);
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
    float n[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float s = (h[i][0] + h[i][1]) + (h[i][2] + h[i][3]);
        s += __shfl_xor_sync(FULL, s, 4);
        s += __shfl_xor_sync(FULL, s, 8);
        s += __shfl_xor_sync(FULL, s, 16);
        const float mean = s * (1.0f / 32.0f);
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            n[i][j] = h[i][j] - mean;
            v = fmaf(n[i][j], n[i][j], v);
        }
        v += __shfl_xor_sync(FULL, v, 4);
        v += __shfl_xor_sync(FULL, v, 8);
        v += __shfl_xor_sync(FULL, v, 16);
        const float rstd = 1.0f / sqrtf(v * (1.0f / 32.0f) + eps);
#pragma unroll
        for (int j = 0; j < 4; j++) n[i][j] = n[i][j] * rstd * wv[j] + bv[j];
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
        *reinterpret_cast<float4*>(aT + (c * 4 + j) * V2G + r * 4) = make_float4(n[0][j], n[1][j], n[2][j], n[3][j]);
}

template <int NCTX, bool CACHE>
__global__ void __launch_bounds__(384, 1) rollout_e32v2_kernel(RolloutParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t bar;
    const Gpt2Layout& L = p.lay;
    const int C = L.n_ctx, NL = L.n_layer;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, W = blockDim.x >> 5;
    const int r = lane & 3, c = lane >> 2;      // lane = r + 4*c: conflict-light transposed tile stores
    float* wsm = sm;
    float* scr = sm + L.total + warp * (32 * V2G);

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)L.total * 4u;
        mbar_expect_tx(&bar, bytes);
        const uint32_t CH = 32768u;
        for (uint32_t off = 0; off < bytes; off += CH)
            bulk_g2s(reinterpret_cast<char*>(wsm) + off, reinterpret_cast<const char*>(p.blob) + off,
                     min(CH, bytes - off), &bar);
    }
    mbar_wait(&bar, 0);

    const int ngroups = (p.B + V2G - 1) / V2G;
    const size_t kv_per_g = (size_t)NL * 2 * C * 32;
    float* kvw = CACHE ? p.kv + (size_t)(blockIdx.x * W + warp) * V2G * kv_per_g : nullptr;
    const float* gw = wsm + L.lnfw;

    for (int grp = blockIdx.x * W + warp; grp < ngroups; grp += gridDim.x * W) {
        const int b0 = grp * V2G + r * 4;            // first trajectory of this lane's tile rows
        float h[4][4];
        {
            const float4 pe = __ldg(reinterpret_cast<const float4*>(p.pe + c * 4));
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (b0 + i < p.B) x = *reinterpret_cast<const float4*>(p.x0 + (size_t)(b0 + i) * p.x0_stride + c * 4);
                h[i][0] = x.x + pe.x; h[i][1] = x.y + pe.y; h[i][2] = x.z + pe.z; h[i][3] = x.w + pe.w;
            }
        }
        for (int s = 0; s < p.S; s++) {
            const int Lk = CACHE ? min(s + 1, C) : 1;
            const int slot = CACHE ? (s % C) : 0;
            for (int l = 0; l < NL; l++) {
                const float* w = wsm + l * L.layer_stride;
                if constexpr (CACHE) {
                    if (p.prefetch && lane < V2G && Lk > 1) {
                        const float* kp = kvw + (size_t)lane * kv_per_g + (size_t)(l * 2) * C * 32;
                        prefetch_l2_bulk(kp, (uint32_t)Lk * 128u);
                        prefetch_l2_bulk(kp + C * 32, (uint32_t)Lk * 128u);
                    }
                }
                tile_ln_to_smem(h, w + L.ln1w, w + L.ln1b, p.eps, scr, r, c);
                __syncwarp();
                if constexpr (CACHE) {
                    float qkv[3][4][4];
#pragma unroll
                    for (int t = 0; t < 3; t++) {
                        const float4 b = lds4(w + L.bqkv + 32 * t + c * 4);
#pragma unroll
                        for (int i = 0; i < 4; i++) { qkv[t][i][0] = b.x; qkv[t][i][1] = b.y; qkv[t][i][2] = b.z; qkv[t][i][3] = b.w; }
                    }
#pragma unroll 2
                    for (int k = 0; k < 32; k++) {
                        const float4 a = lds4(scr + k * V2G + r * 4);
                        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                        for (int t = 0; t < 3; t++) {
                            const float4 wq = lds4(w + L.wqkv + k * 96 + 32 * t + c * 4);
                            const float wv[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
                            for (int i = 0; i < 4; i++)
#pragma unroll
                                for (int j = 0; j < 4; j++) qkv[t][i][j] = fmaf(av[i], wv[j], qkv[t][i][j]);
                        }
                    }
                    // append k, v rows (16 bytes per lane, 128 B per trajectory row per instruction)
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        float* kp = kvw + (size_t)(r * 4 + i) * kv_per_g + (size_t)(l * 2) * C * 32 + slot * 32 + c * 4;
                        *reinterpret_cast<float4*>(kp) = make_float4(qkv[1][i][0], qkv[1][i][1], qkv[1][i][2], qkv[1][i][3]);
                        *reinterpret_cast<float4*>(kp + C * 32) = make_float4(qkv[2][i][0], qkv[2][i][1], qkv[2][i][2], qkv[2][i][3]);
                    }
                    __syncwarp();                 // LN tile consumed; ring rows visible to the warp
#pragma unroll
                    for (int i = 0; i < 4; i++)     // queries, row-major per trajectory: qs[g][32]
                        *reinterpret_cast<float4*>(scr + (r * 4 + i) * 32 + c * 4) =
                            make_float4(qkv[0][i][0], qkv[0][i][1], qkv[0][i][2], qkv[0][i][3]);
                    __syncwarp();
                    float o[2][8];
#pragma unroll
                    for (int pp = 0; pp < 2; pp++) {
                        const int id = lane + 32 * pp, g = id >> 2, hh = id & 3;
                        float q[8];
                        const float4 qa = lds4(scr + g * 32 + hh * 8), qb = lds4(scr + g * 32 + hh * 8 + 4);
                        q[0] = qa.x; q[1] = qa.y; q[2] = qa.z; q[3] = qa.w; q[4] = qb.x; q[5] = qb.y; q[6] = qb.z; q[7] = qb.w;
                        const float* Kp = kvw + (size_t)g * kv_per_g + (size_t)(l * 2) * C * 32 + hh * 8;
                        lane_attention<NCTX>(q, Kp, Kp + C * 32, Lk, o[pp]);
                    }
                    __syncwarp();                 // every lane has read its queries
#pragma unroll
                    for (int pp = 0; pp < 2; pp++) {
                        const int id = lane + 32 * pp, g = id >> 2, hh = id & 3;
#pragma unroll
                        for (int d = 0; d < 8; d++) scr[(hh * 8 + d) * V2G + g] = o[pp][d];
                    }
                    __syncwarp();
                } else {
                    float vv[4][4];
                    {
                        const float4 b = lds4(w + L.bqkv + 64 + c * 4);
#pragma unroll
                        for (int i = 0; i < 4; i++) { vv[i][0] = b.x; vv[i][1] = b.y; vv[i][2] = b.z; vv[i][3] = b.w; }
                    }
                    tile_gemv<1>(scr, w + L.wqkv + 64 + c * 4, 32, 96, r, vv);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        *reinterpret_cast<float4*>(scr + (c * 4 + j) * V2G + r * 4) = make_float4(vv[0][j], vv[1][j], vv[2][j], vv[3][j]);
                    __syncwarp();
                }
                {   // output projection + residual
                    float acc[4][4];
                    const float4 b = lds4(w + L.bproj + c * 4);
#pragma unroll
                    for (int i = 0; i < 4; i++) { acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.z; acc[i][3] = b.w; }
                    tile_gemv<1>(scr, w + L.wproj + c * 4, 32, 32, r, acc);
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) h[i][j] += acc[i][j];
                }
                __syncwarp();
                tile_ln_to_smem(h, w + L.ln2w, w + L.ln2b, p.eps, scr, r, c);
                __syncwarp();
                float f[4][16];
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const float4 b = lds4(w + L.bfc + 32 * m + c * 4);      // unit 32*m + 4*c + j
#pragma unroll
                    for (int i = 0; i < 4; i++) { f[i][4 * m] = b.x; f[i][4 * m + 1] = b.y; f[i][4 * m + 2] = b.z; f[i][4 * m + 3] = b.w; }
                }
                tile_gemv<4, 32>(scr, w + L.wfc + c * 4, 32, 128, r, f);
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 16; j++) f[i][j] = gelu_new_sig(f[i][j]);
                float acc[4][4];
                {
                    const float4 b = lds4(w + L.bpr + c * 4);
#pragma unroll
                    for (int i = 0; i < 4; i++) { acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.z; acc[i][3] = b.w; }
                }
#pragma unroll
                for (int qd = 0; qd < 4; qd++) {
                    // hidden units 32*qd .. 32*qd+31: this lane holds units 32*qd + 4*c + i2 in f[.][4*qd + i2]
                    __syncwarp();
#pragma unroll
                    for (int i2 = 0; i2 < 4; i2++)
                        *reinterpret_cast<float4*>(scr + (4 * c + i2) * V2G + r * 4) =
                            make_float4(f[0][4 * qd + i2], f[1][4 * qd + i2], f[2][4 * qd + i2], f[3][4 * qd + i2]);
                    __syncwarp();
                    tile_gemv<1>(scr, w + L.wpr + (32 * qd) * 32 + c * 4, 32, 32, r, acc);
                }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) h[i][j] += acc[i][j];
                __syncwarp();
            }
            // final LayerNorm + linear head
            tile_ln_to_smem(h, gw, gw + 32, p.eps, scr, r, c);
            __syncwarp();
            float y[4][4];
            {
                const float4 b = lds4(gw + 64 + 1024 + c * 4);
#pragma unroll
                for (int i = 0; i < 4; i++) { y[i][0] = b.x; y[i][1] = b.y; y[i][2] = b.z; y[i][3] = b.w; }
            }
            tile_gemv<1>(scr, gw + 64 + c * 4, 32, 32, r, y);
            const int pnext = CACHE ? min(s + 1, C - 1) : 0;
            const float4 pe = __ldg(reinterpret_cast<const float4*>(p.pe + pnext * 32 + c * 4));
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (b0 + i < p.B)
                    *reinterpret_cast<float4*>(p.out + (size_t)(b0 + i) * p.out_stride + (size_t)s * 32 + c * 4) =
                        make_float4(y[i][0], y[i][1], y[i][2], y[i][3]);
                h[i][0] = y[i][0] + pe.x; h[i][1] = y[i][1] + pe.y; h[i][2] = y[i][2] + pe.z; h[i][3] = y[i][3] + pe.w;
            }
            __syncwarp();
        }
        if (CACHE && p.presents != nullptr) {
            const int Tk = min(p.S, C);
            const int H = 4, hd = 8;
            const size_t per_l = (size_t)2 * p.B * H * Tk * hd;
            const int hh = lane >> 3, d = lane & 7;
            for (int g = 0; g < V2G; g++) {
                const int b = grp * V2G + g;
                if (b >= p.B) break;
                for (int rr = 0; rr < NL * 2 * Tk; rr++) {
                    const int t = rr % Tk, lw = rr / Tk;
                    const int which = lw & 1, l = lw >> 1;
                    const int sl = (p.S - Tk + t) % C;
                    p.presents[(size_t)l * per_l + ((((size_t)which * p.B + b) * H + hh) * Tk + t) * hd + d] =
                        kvw[(size_t)g * kv_per_g + ((size_t)lw * C + sl) * 32 + lane];
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1 (E = 32, head size 8), tensor-core generation: 3xTF32 on mma.sync.m16n8k8 (SASS HMMA.1688.F32.TF32).
// Why: ncu shows the SIMT GEMVs bound by shared-memory return bandwidth at ~40 % of the FP32 pipe; the tensor
// path measured on this B200 (scripts/probes/mma_probe.cu) sustains 278 TFLOP/s TF32 with 4 warps per SM, i.e.
// ~93 TFLOP/s for the error-compensated 3xTF32 product (a_lo*w_hi + a_hi*w_lo + a_hi*w_hi, FP32 accumulate),
// which keeps the FP32 parity bar (plain TF32 would not).  tcgen05 needs M >= 64 rows per CTA tile, TMEM
// round trips and hi/lo weight planes (2 x 203 KB > shared memory); at 16 trajectories per warp the legacy
// warp-level MMA keeps the warp-autonomous, weights-resident structure of the SIMT kernels.
//   * one warp = one 16-trajectory M-tile; lane (g = lane>>2, q = lane&3) holds rows g and g+8;
//   * the residual stream lives in C-fragment layout; a C fragment of n-tile t IS the A fragment of k-step t
//     when the K order inside a k-step is taken as (8t+0, 8t+2, 8t+4, 8t+6 | 8t+1, 8t+3, 8t+5, 8t+7) — so
//     activations go GEMV -> GEMV through registers only (LayerNorm: 2 xor-shuffles over q);
//   * B fragments come from an XOR-swizzled FP32 weight image in shared memory (conflict-free), split into
//     TF32 hi/lo on the fly (cvt.rna + subtract);
//   * the MLP is fused per 8 hidden units: MLP-in tile -> gelu -> A fragment -> 4 MLP-out tiles.
// ------------------------------------------------------------------------------------------------
constexpr int V4G = 16;
constexpr int V4QS = 36;          // row stride (floats) of the per-warp q / attention-output staging tile

__device__ __forceinline__ uint32_t tf32_hi(float x) {
    return tf32_rna_bits(x);
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_hi(x);
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += A(16x8) * W(8x8) with error compensation; w0/w1 = the lane's two FP32 weights of the tile
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float w0,
                                     float w1) {
    uint32_t bh0, bl0, bh1, bl1;
    split_tf32(w0, bh0, bl0);
    split_tf32(w1, bh1, bl1);
    mma_tf32(c, al, bh0, bh1);
    mma_tf32(c, ah, bl0, bl1);
    mma_tf32(c, ah, bh0, bh1);
}
// C-fragment values (row g: x0, x1 ; row g+8: x2, x3) -> A fragment (a0,a1,a2,a3) = (x0, x2, x1, x3), hi/lo
__device__ __forceinline__ void frag_c2a(const float (&x)[4], uint32_t (&ah)[4], uint32_t (&al)[4]) {
    split_tf32(x[0], ah[0], al[0]);
    split_tf32(x[2], ah[1], al[1]);
    split_tf32(x[1], ah[2], al[2]);
    split_tf32(x[3], ah[3], al[3]);
}
// swizzled weight element (k, n) of a [K][N] matrix: k*N + (n ^ (((k >> 1) & 3) << 3)).  For the B fragment of
// k-step s / n-tile t the lane reads rows 8s+2q and 8s+2q+1 at physical column 8*(t^q) + g.
__device__ __forceinline__ void ldb(const float* W, int N, int s, int t, int g, int q, float& w0, float& w1) {
    const float* p = W + (8 * s + 2 * q) * N + 8 * (t ^ q) + g;
    w0 = p[0];
    w1 = p[N];
}
// acc[t] (t < NT) += X * W[:, 8*(t0+t) .. +8) for X given as KS k-step fragments
template <int KS, int NT>
__device__ __forceinline__ void gemv_mma(float (&acc)[NT][4], const uint32_t (&ah)[KS][4], const uint32_t (&al)[KS][4],
                                         const float* W, int N, int t0, int g, int q) {
#pragma unroll
    for (int s = 0; s < KS; s++)
#pragma unroll
        for (int t = 0; t < NT; t++) {       // NT independent accumulator chains in flight
            float w0, w1;
            ldb(W, N, s, t0 + t, g, q, w0, w1);
            mma3(acc[t], ah[s], al[s], w0, w1);
        }
}
// LayerNorm over the 32 channels of rows g and g+8 (fragment layout), result as A fragments
__device__ __forceinline__ void frag_ln(const float (&h)[4][4], const float* lw, const float* lb, float eps, int q,
                                        uint32_t (&ah)[4][4], uint32_t (&al)[4][4]) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int t = 0; t < 4; t++) { s0 += h[t][0] + h[t][1]; s1 += h[t][2] + h[t][3]; }
    s0 += __shfl_xor_sync(FULL, s0, 1); s0 += __shfl_xor_sync(FULL, s0, 2);
    s1 += __shfl_xor_sync(FULL, s1, 1); s1 += __shfl_xor_sync(FULL, s1, 2);
    const float m0 = s0 * (1.0f / 32.0f), m1 = s1 * (1.0f / 32.0f);
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const float a = h[t][0] - m0, b = h[t][1] - m0, c = h[t][2] - m1, d = h[t][3] - m1;
        v0 = fmaf(a, a, v0); v0 = fmaf(b, b, v0);
        v1 = fmaf(c, c, v1); v1 = fmaf(d, d, v1);
    }
    v0 += __shfl_xor_sync(FULL, v0, 1); v0 += __shfl_xor_sync(FULL, v0, 2);
    v1 += __shfl_xor_sync(FULL, v1, 1); v1 += __shfl_xor_sync(FULL, v1, 2);
    const float r0 = 1.0f / sqrtf(v0 * (1.0f / 32.0f) + eps), r1 = 1.0f / sqrtf(v1 * (1.0f / 32.0f) + eps);
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const float2 w2 = *reinterpret_cast<const float2*>(lw + 8 * t + 2 * q);
        const float2 b2 = *reinterpret_cast<const float2*>(lb + 8 * t + 2 * q);
        float x[4];
        x[0] = (h[t][0] - m0) * r0 * w2.x + b2.x;
        x[1] = (h[t][1] - m0) * r0 * w2.y + b2.y;
        x[2] = (h[t][2] - m1) * r1 * w2.x + b2.x;
        x[3] = (h[t][3] - m1) * r1 * w2.y + b2.y;
        frag_c2a(x, ah[t], al[t]);
    }
}
// same LayerNorm, result stored as FP32 rows (row g -> r0, row g + 8 -> r1; the lane's columns 8t + 2q, + 1)
__device__ __forceinline__ void frag_ln_store(const float (&h)[4][4], const float* lw, const float* lb, float eps, int q,
                                              float* r0, float* r1) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int t = 0; t < 4; t++) { s0 += h[t][0] + h[t][1]; s1 += h[t][2] + h[t][3]; }
    s0 += __shfl_xor_sync(FULL, s0, 1); s0 += __shfl_xor_sync(FULL, s0, 2);
    s1 += __shfl_xor_sync(FULL, s1, 1); s1 += __shfl_xor_sync(FULL, s1, 2);
    const float m0 = s0 * (1.0f / 32.0f), m1 = s1 * (1.0f / 32.0f);
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const float a = h[t][0] - m0, b = h[t][1] - m0, c = h[t][2] - m1, d = h[t][3] - m1;
        v0 = fmaf(a, a, v0); v0 = fmaf(b, b, v0);
        v1 = fmaf(c, c, v1); v1 = fmaf(d, d, v1);
    }
    v0 += __shfl_xor_sync(FULL, v0, 1); v0 += __shfl_xor_sync(FULL, v0, 2);
    v1 += __shfl_xor_sync(FULL, v1, 1); v1 += __shfl_xor_sync(FULL, v1, 2);
    const float i0 = 1.0f / sqrtf(v0 * (1.0f / 32.0f) + eps), i1 = 1.0f / sqrtf(v1 * (1.0f / 32.0f) + eps);
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const float2 w2 = *reinterpret_cast<const float2*>(lw + 8 * t + 2 * q);
        const float2 b2 = *reinterpret_cast<const float2*>(lb + 8 * t + 2 * q);
        *reinterpret_cast<float2*>(r0 + 8 * t + 2 * q) =
            make_float2((h[t][0] - m0) * i0 * w2.x + b2.x, (h[t][1] - m0) * i0 * w2.y + b2.y);
        *reinterpret_cast<float2*>(r1 + 8 * t + 2 * q) =
            make_float2((h[t][2] - m1) * i1 * w2.x + b2.x, (h[t][3] - m1) * i1 * w2.y + b2.y);
    }
}
__device__ __forceinline__ void frag_bias(float (&c)[4], const float* b, int t, int q) {
    const float2 b2 = *reinterpret_cast<const float2*>(b + 8 * t + 2 * q);
    c[0] = b2.x; c[1] = b2.y; c[2] = b2.x; c[3] = b2.y;
}

template <int NCTX, bool CACHE>
__global__ void __launch_bounds__(CACHE ? 256 : 384, 1) rollout_e32v4_kernel(RolloutParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t bar;
    const Gpt2Layout& L = p.lay;
    const int C = L.n_ctx, NL = L.n_layer;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, W = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    float* wsm = sm;
    float* scr = sm + L.total + warp * (V4G * V4QS);      // only used (and allocated) with the KV window

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)L.total * 4u;
        mbar_expect_tx(&bar, bytes);
        const uint32_t CH = 32768u;
        for (uint32_t off = 0; off < bytes; off += CH)
            bulk_g2s(reinterpret_cast<char*>(wsm) + off, reinterpret_cast<const char*>(p.blob_sw) + off,
                     min(CH, bytes - off), &bar);
    }
    mbar_wait(&bar, 0);

    const int ngroups = (p.B + V4G - 1) / V4G;
    const size_t kv_per_g = (size_t)NL * 2 * C * 32;
    float* kvw = CACHE ? p.kv + (size_t)(blockIdx.x * W + warp) * V4G * kv_per_g : nullptr;
    const float* gw = wsm + L.lnfw;

    for (int grp = blockIdx.x * W + warp; grp < ngroups; grp += gridDim.x * W) {
        const int br[2] = {grp * V4G + g, grp * V4G + g + 8};     // the lane's two trajectories
        float h[4][4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const float2 pe = __ldg(reinterpret_cast<const float2*>(p.pe + 8 * t + 2 * q));
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                float2 x = make_float2(0.f, 0.f);
                if (br[rr] < p.B) x = *reinterpret_cast<const float2*>(p.x0 + (size_t)br[rr] * p.x0_stride + 8 * t + 2 * q);
                h[t][2 * rr] = x.x + pe.x;
                h[t][2 * rr + 1] = x.y + pe.y;
            }
        }
        for (int s = 0; s < p.S; s++) {
            const int Lk = CACHE ? min(s + 1, C) : 1;
            const int slot = CACHE ? (s % C) : 0;
            for (int l = 0; l < NL; l++) {
                const float* w = wsm + l * L.layer_stride;
                if constexpr (CACHE) {
                    if ((p.prefetch & 1) && lane < (NCTX <= 32 ? 8 : 4) && Lk > 1) {      // windows of the first attention pass
                        const float* kp = kvw + (size_t)lane * kv_per_g + (size_t)(l * 2) * C * 32;
                        prefetch_l2_bulk(kp, (uint32_t)Lk * 128u);
                        prefetch_l2_bulk(kp + C * 32, (uint32_t)Lk * 128u);
                    }
                }
                uint32_t ah[4][4], al[4][4];
                frag_ln(h, w + L.ln1w, w + L.ln1b, p.eps, q, ah, al);
                float o[4][4];                    // attention output, fragment layout
                if constexpr (CACHE) {
                    float qf[4][4], kf[4][4], vf[4][4];
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        frag_bias(qf[t], w + L.bqkv, t, q);
                        frag_bias(kf[t], w + L.bqkv + 32, t, q);
                        frag_bias(vf[t], w + L.bqkv + 64, t, q);
                    }
                    gemv_mma<4, 4>(qf, ah, al, w + L.wqkv, 96, 0, g, q);
                    gemv_mma<4, 4>(kf, ah, al, w + L.wqkv, 96, 4, g, q);
                    gemv_mma<4, 4>(vf, ah, al, w + L.wqkv, 96, 8, g, q);
                    // append k, v rows of the two trajectories; stage q row-major for the per-pair attention
#pragma unroll
                    for (int rr = 0; rr < 2; rr++) {
                        float* kp = kvw + (size_t)(g + 8 * rr) * kv_per_g + (size_t)(l * 2) * C * 32 + slot * 32 + 2 * q;
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            *reinterpret_cast<float2*>(kp + 8 * t) = make_float2(kf[t][2 * rr], kf[t][2 * rr + 1]);
                            *reinterpret_cast<float2*>(kp + C * 32 + 8 * t) = make_float2(vf[t][2 * rr], vf[t][2 * rr + 1]);
                            *reinterpret_cast<float2*>(scr + (g + 8 * rr) * V4QS + 8 * t + 2 * q) =
                                make_float2(qf[t][2 * rr], qf[t][2 * rr + 1]);
                        }
                    }
                    __syncwarp();
                    // attention in 4 passes of 4 trajectories: two lanes per (trajectory, head) pair, 16 window rows in
                    // flight per lane (the scheme of rollout_e32_kernel<4>); the windows of the NEXT pass are pulled
                    // into L2 while this pass computes, so only ~32 KB per warp have to survive in L2
                    if constexpr (NCTX <= 32) {
                        // two passes of 8 trajectories: a lane owns a (trajectory, head) pair outright (the scheme of
                        // rollout_e32_kernel<8>), 16 window rows in flight per lane
#pragma unroll 1
                        for (int t8 = 0; t8 < 2; t8++) {
                            const float* kvl = kvw + (size_t)(8 * t8) * kv_per_g + (size_t)(l * 2) * C * 32;
                            if ((p.prefetch & 2) && t8 < 1 && lane < 8 && Lk > 1) {
                                const float* kp = kvl + (size_t)(8 + lane) * kv_per_g;
                                prefetch_l2_bulk(kp, (uint32_t)Lk * 128u);
                                prefetch_l2_bulk(kp + C * 32, (uint32_t)Lk * 128u);
                            }
                            float* tile = scr + 8 * t8 * V4QS;
                            if (Lk == NCTX)
                                warp_attn_e32<8, NCTX, true, V4QS, V4QS>(tile, kvl, kv_per_g, C, Lk, tile, lane);
                            else
                                warp_attn_e32<8, NCTX, false, V4QS, V4QS>(tile, kvl, kv_per_g, C, Lk, tile, lane);
                        }
                    } else
#pragma unroll 1
                    for (int t4 = 0; t4 < 4; t4++) {
                        const float* kvl = kvw + (size_t)(4 * t4) * kv_per_g + (size_t)(l * 2) * C * 32;
                        if ((p.prefetch & 2) && t4 < 3 && lane < 4 && Lk > 1) {
                            const float* kp = kvl + (size_t)(4 + lane) * kv_per_g;
                            prefetch_l2_bulk(kp, (uint32_t)Lk * 128u);
                            prefetch_l2_bulk(kp + C * 32, (uint32_t)Lk * 128u);
                        }
                        float* tile = scr + 4 * t4 * V4QS;
                        if (Lk == NCTX)
                            warp_attn_e32<4, NCTX, true, V4QS, V4QS>(tile, kvl, kv_per_g, C, Lk, tile, lane);
                        else
                            warp_attn_e32<4, NCTX, false, V4QS, V4QS>(tile, kvl, kv_per_g, C, Lk, tile, lane);
                    }
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < 4; t++)
#pragma unroll
                        for (int rr = 0; rr < 2; rr++) {
                            const float2 v2 = *reinterpret_cast<const float2*>(scr + (g + 8 * rr) * V4QS + 8 * t + 2 * q);
                            o[t][2 * rr] = v2.x;
                            o[t][2 * rr + 1] = v2.y;
                        }
                    __syncwarp();
                } else {
                    // window of one key: the head output is v itself
#pragma unroll
                    for (int t = 0; t < 4; t++) frag_bias(o[t], w + L.bqkv + 64, t, q);
                    gemv_mma<4, 4>(o, ah, al, w + L.wqkv, 96, 8, g, q);
                }
                // ---- output projection + residual ----
#pragma unroll
                for (int t = 0; t < 4; t++) frag_c2a(o[t], ah[t], al[t]);
                {
                    float acc[4][4];
#pragma unroll
                    for (int t = 0; t < 4; t++) frag_bias(acc[t], w + L.bproj, t, q);
                    gemv_mma<4, 4>(acc, ah, al, w + L.wproj, 32, 0, g, q);
#pragma unroll
                    for (int t = 0; t < 4; t++)
#pragma unroll
                        for (int i = 0; i < 4; i++) h[t][i] += acc[t][i];
                }
                // ---- LN2 + MLP fused per 8 hidden units ----
                frag_ln(h, w + L.ln2w, w + L.ln2b, p.eps, q, ah, al);
                float pr[4][4];
#pragma unroll
                for (int t = 0; t < 4; t++) frag_bias(pr[t], w + L.bpr, t, q);
#pragma unroll 1
                for (int tf0 = 0; tf0 < 16; tf0 += 4) {
                    // four MLP-in tiles (32 hidden units) at a time: 4 independent MMA chains
                    float fc[4][4];
#pragma unroll
                    for (int j = 0; j < 4; j++) frag_bias(fc[j], w + L.bfc, tf0 + j, q);
                    gemv_mma<4, 4>(fc, ah, al, w + L.wfc, 128, tf0, g, q);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
#pragma unroll
                        for (int i = 0; i < 4; i++) fc[j][i] = gelu_new_sig(fc[j][i]);
                        uint32_t fh[4], fl[4];
                        frag_c2a(fc[j], fh, fl);
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            float w0, w1;
                            ldb(w + L.wpr, 32, tf0 + j, u, g, q, w0, w1);
                            mma3(pr[u], fh, fl, w0, w1);
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < 4; t++)
#pragma unroll
                    for (int i = 0; i < 4; i++) h[t][i] += pr[t][i];
            }
            // ---- final LayerNorm + linear head ----
            uint32_t ah[4][4], al[4][4];
            frag_ln(h, gw, gw + 32, p.eps, q, ah, al);
            float y[4][4];
#pragma unroll
            for (int t = 0; t < 4; t++) frag_bias(y[t], gw + 64 + 1024, t, q);
            gemv_mma<4, 4>(y, ah, al, gw + 64, 32, 0, g, q);
            const int pnext = CACHE ? min(s + 1, C - 1) : 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const float2 pe = __ldg(reinterpret_cast<const float2*>(p.pe + pnext * 32 + 8 * t + 2 * q));
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    if (br[rr] < p.B)
                        *reinterpret_cast<float2*>(p.out + (size_t)br[rr] * p.out_stride + (size_t)s * 32 + 8 * t + 2 * q) =
                            make_float2(y[t][2 * rr], y[t][2 * rr + 1]);
                    h[t][2 * rr] = y[t][2 * rr] + pe.x;
                    h[t][2 * rr + 1] = y[t][2 * rr + 1] + pe.y;
                }
            }
        }
        if (CACHE && p.presents != nullptr) {
            __syncwarp();
            const int Tk = min(p.S, C);
            const int H = 4, hd = 8;
            const size_t per_l = (size_t)2 * p.B * H * Tk * hd;
            const int hh = lane >> 3, d = lane & 7;
            for (int tg = 0; tg < V4G; tg++) {
                const int b = grp * V4G + tg;
                if (b >= p.B) break;
                for (int rr = 0; rr < NL * 2 * Tk; rr++) {
                    const int t = rr % Tk, lw = rr / Tk;
                    const int which = lw & 1, l = lw >> 1;
                    const int sl = (p.S - Tk + t) % C;
                    p.presents[(size_t)l * per_l + ((((size_t)which * p.B + b) * H + hh) * Tk + t) * hd + d] =
                        kvw[(size_t)tg * kv_per_g + ((size_t)lw * C + sl) * 32 + lane];
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1w (E = 32, KV window): WARP-SPECIALISED streaming kernel (rollout variant 11).
// Measured on the kernels above (profiles/README.md, round 2): a warp that does both the dense layer arithmetic and
// the window attention alternates between a phase that needs no memory and a phase that only waits for memory; with
// the 8 warps the register file allows, on average 3.5 warps have window rows in flight and the HBM stream runs at
// 5.1-5.5 TB/s; the tensor-core dense part needs only 4 warps per SM (rollout_e32v4_kernel without the window:
// 35 ms at 4 warps, 31 ms at 8-12).  Here the two phases run in DIFFERENT warps of a persistent CTA:
//   * 4 dense warps (one per scheduler): 3xTF32 mma.sync tile arithmetic of rollout_e32v4_kernel for TWO
//     16-trajectory groups each, alternately (residual streams of both groups in registers);
//   * 8 attention warps, one per group in flight (8 groups = 128 trajectories per SM): warp_attn_e32<4> over the
//     group's window, 16 rows in flight per lane, always streaming;
//   * hand-off per (group, step, layer) through a 16 x 32 query / attention-output tile in shared memory and two
//     mbarriers (qready: dense -> attention after k, v were appended to the ring and q staged; oready: back);
//     a group's dense warp and attention warp sit on the same scheduler (warp ids d, 4 + d, 8 + d).
// Arithmetic (and therefore every result bit) is the one of rollout_e32v4_kernel<NCTX, true>.
// ------------------------------------------------------------------------------------------------
constexpr int WS_DW = 4;              // dense warps
constexpr int WS_NG = 8;              // groups in flight per CTA = attention warps
constexpr int WS_THREADS = 32 * (WS_DW + WS_NG);
#ifndef WS_PROF
#define WS_PROF 0          // 1: per-warp wait / phase cycle counters (TRPX_WS_PROF=1 prints them)
#endif

template <int NCTX>
__global__ void __launch_bounds__(WS_THREADS, 1) rollout_e32ws_kernel(RolloutParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t qready[WS_NG];
    __shared__ __align__(8) uint64_t oready[WS_NG];
    const Gpt2Layout& L = p.lay;
    const int C = L.n_ctx, NL = L.n_layer;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* wsm = sm;
    float* tiles = sm + L.total;                       // [WS_NG][V4G][V4QS]

    if (tid == 0) {
        mbar_init(&bar, 1);
        for (int i = 0; i < WS_NG; i++) { mbar_init(&qready[i], 1); mbar_init(&oready[i], 1); }
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)L.total * 4u;
        mbar_expect_tx(&bar, bytes);
        const uint32_t CH = 32768u;
        for (uint32_t off = 0; off < bytes; off += CH)
            bulk_g2s(reinterpret_cast<char*>(wsm) + off, reinterpret_cast<const char*>(p.blob_sw) + off,
                     min(CH, bytes - off), &bar);
    }
    mbar_wait(&bar, 0);

    const int ngroups = (p.B + V4G - 1) / V4G;
    const size_t kv_per_g = (size_t)NL * 2 * C * 32;
    const int nslots = gridDim.x * WS_NG;               // groups in flight on the whole device
    // L2 prefetch of a group's layer window (4 attention passes of 4 trajectories): the dense warp requests the first
    // pfD passes when it starts the stage that ends with the hand-over, the attention warp stays pfA passes ahead
    const int pfD = p.prefetch & 15, pfA = (p.prefetch >> 4) & 15;
    // slot gi of CTA b works on groups  r * nslots + gi * gridDim.x + b  (r = 0, 1, ...): a partly filled last round
    // spreads over all SMs

    if (warp >= WS_DW) {
        // ================= attention warp: one slot =================
        const int gi = warp - WS_DW;
        float* tile = tiles + gi * (V4G * V4QS);
        float* kvs = p.kv + (size_t)(blockIdx.x * WS_NG + gi) * V4G * kv_per_g;
        uint32_t ph = 0;
#if WS_PROF
        long long twait = 0;
        const long long tstart = clock64();
#endif
        for (int grp = gi * gridDim.x + blockIdx.x; grp < ngroups; grp += nslots) {
            for (int s = 0; s < p.S; s++) {
                const int Lk = min(s + 1, C);
                for (int l = 0; l < NL; l++) {
#if WS_PROF
                    const long long tw0 = clock64();
#endif
                    mbar_wait_backoff(&qready[gi], ph, p.spin_ns);
                    ph ^= 1u;
#if WS_PROF
                    twait += clock64() - tw0;
#endif
                    int pfdone = pfD;                  // passes whose windows have been requested into L2
                    if (pfA > 0 && Lk > 1 && pfdone < pfA) {
                        const int upto = min(pfA, 4) - 1;
                        if (lane < 4 * (upto - pfdone + 1)) {
                            const float* kp = kvs + (size_t)(4 * pfdone + lane) * kv_per_g + (size_t)(l * 2) * C * 32;
                            prefetch_l2_bulk(kp, (uint32_t)Lk * 128u);
                            prefetch_l2_bulk(kp + C * 32, (uint32_t)Lk * 128u);
                        }
                        pfdone = upto + 1;
                    }
                    {
                        // ---- q, k, v of the group's 16 trajectories from LN1(h) (3xTF32 tile products) ----
                        const float* w = wsm + l * L.layer_stride;
                        const int g = lane >> 2, q = lane & 3;
                        const int slot = s % C;
                        uint32_t ah[4][4], al[4][4];
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            float x[4];
#pragma unroll
                            for (int rr = 0; rr < 2; rr++) {
                                const float2 v2 = *reinterpret_cast<const float2*>(tile + (g + 8 * rr) * V4QS + 8 * t + 2 * q);
                                x[2 * rr] = v2.x;
                                x[2 * rr + 1] = v2.y;
                            }
                            frag_c2a(x, ah[t], al[t]);
                        }
                        __syncwarp();                          // every lane holds its fragments: the tile may take q
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            float f[4][4];
#pragma unroll
                            for (int t = 0; t < 4; t++) frag_bias(f[t], w + L.bqkv + 32 * c, t, q);
                            gemv_mma<4, 4>(f, ah, al, w + L.wqkv, 96, 4 * c, g, q);
#pragma unroll
                            for (int rr = 0; rr < 2; rr++) {
                                float* dst = (c == 0) ? tile + (g + 8 * rr) * V4QS + 2 * q
                                                      : kvs + (size_t)(g + 8 * rr) * kv_per_g +
                                                            (size_t)(l * 2 + (c - 1)) * C * 32 + slot * 32 + 2 * q;
#pragma unroll
                                for (int t = 0; t < 4; t++)
                                    *reinterpret_cast<float2*>(dst + 8 * t) = make_float2(f[t][2 * rr], f[t][2 * rr + 1]);
                            }
                        }
                        __syncwarp();                          // q tile and ring rows visible to the whole warp
                    }
#pragma unroll 1
                    for (int t4 = (WS_PROF && (p.prefetch & 256)) ? 4 : 0; t4 < 4; t4++) {
                        const float* kvl = kvs + (size_t)(4 * t4) * kv_per_g + (size_t)(l * 2) * C * 32;
                        if (pfA > 0 && Lk > 1) {
                            const int upto = min(t4 + pfA, 3);
                            if (pfdone <= upto) {
                                const int np = 4 * (upto - pfdone + 1);
                                if (lane < np) {
                                    const float* kp = kvs + (size_t)(4 * pfdone + lane) * kv_per_g + (size_t)(l * 2) * C * 32;
                                    prefetch_l2_bulk(kp, (uint32_t)Lk * 128u);
                                    prefetch_l2_bulk(kp + C * 32, (uint32_t)Lk * 128u);
                                }
                                pfdone = upto + 1;
                            }
                        }
                        float* tl = tile + 4 * t4 * V4QS;
                        if (Lk == NCTX)
                            warp_attn_e32<4, NCTX, true, V4QS, V4QS>(tl, kvl, kv_per_g, C, Lk, tl, lane);
                        else
                            warp_attn_e32<4, NCTX, false, V4QS, V4QS>(tl, kvl, kv_per_g, C, Lk, tl, lane);
                    }
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&oready[gi]);
                }
            }
        }
#if WS_PROF
        if (p.prof != nullptr && blockIdx.x == 0 && lane == 0) {
            p.prof[2 * warp] = twait;
            p.prof[2 * warp + 1] = clock64() - tstart;
        }
#endif
        return;
    }

    // ================= dense warp: slots `warp` (A) and `warp + 4` (B), alternately =================
    const int g = lane >> 2, q = lane & 3;
    const float* gw = wsm + L.lnfw;
    const int nsteps = p.S * NL;
    float h[4][4], hB[4][4];           // residual stream of the slot being worked on / of the other slot
    uint32_t ph = 0, phB = 0;
#if WS_PROF
    long long twait = 0, tA = 0, tB = 0, tC = 0;
    const long long tstart = clock64();
#endif
    for (int r = 0;; r++) {
        const int grp0 = r * nslots + warp * gridDim.x + blockIdx.x;
        const int grp1 = grp0 + WS_DW * gridDim.x;
        if (grp0 >= ngroups) break;
        const bool act1 = grp1 < ngroups;
        for (int n = 0; n <= nsteps; n++) {
            const int s = n / NL, l = n - s * NL;
#pragma unroll 1
            for (int k = 0; k < 2; k++) {
                if (k == 0 || act1) {
                    const int gi = warp + WS_DW * k;
                    const int grp = k ? grp1 : grp0;
                    float* tile = tiles + gi * (V4G * V4QS);
                    float* kvs = p.kv + (size_t)(blockIdx.x * WS_NG + gi) * V4G * kv_per_g;
                    const int br[2] = {grp * V4G + g, grp * V4G + g + 8};     // the lane's two trajectories
                    if (n < nsteps && lane < 4 * pfD && s > 0) {      // windows of the first pfD attention passes of (s, l)
                        const float* kp = kvs + (size_t)lane * kv_per_g + (size_t)(l * 2) * C * 32;
                        const uint32_t nb = (uint32_t)min(s + 1, C) * 128u;
                        prefetch_l2_bulk(kp, nb);
                        prefetch_l2_bulk(kp + C * 32, nb);
                    }
                    if (n == 0) {
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            const float2 pe = __ldg(reinterpret_cast<const float2*>(p.pe + 8 * t + 2 * q));
#pragma unroll
                            for (int rr = 0; rr < 2; rr++) {
                                float2 x = make_float2(0.f, 0.f);
                                if (br[rr] < p.B)
                                    x = *reinterpret_cast<const float2*>(p.x0 + (size_t)br[rr] * p.x0_stride + 8 * t + 2 * q);
                                h[t][2 * rr] = x.x + pe.x;
                                h[t][2 * rr + 1] = x.y + pe.y;
                            }
                        }
                    } else {
                        // ---- second half of the layer whose attention has just been done ----
                        const float* w = wsm + (l == 0 ? NL - 1 : l - 1) * L.layer_stride;
#if WS_PROF
                        const long long tw0 = clock64();
#endif
                        mbar_wait_backoff(&oready[gi], ph, p.spin_ns);
                        ph ^= 1u;
#if WS_PROF
                        const long long tp1 = clock64();
                        twait += tp1 - tw0;
#endif
                        uint32_t ah[4][4], al[4][4];
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            float o[4];
#pragma unroll
                            for (int rr = 0; rr < 2; rr++) {
                                const float2 v2 = *reinterpret_cast<const float2*>(tile + (g + 8 * rr) * V4QS + 8 * t + 2 * q);
                                o[2 * rr] = v2.x;
                                o[2 * rr + 1] = v2.y;
                            }
                            frag_c2a(o, ah[t], al[t]);
                        }
                        {
                            float acc[4][4];
#pragma unroll
                            for (int t = 0; t < 4; t++) frag_bias(acc[t], w + L.bproj, t, q);
                            gemv_mma<4, 4>(acc, ah, al, w + L.wproj, 32, 0, g, q);
#pragma unroll
                            for (int t = 0; t < 4; t++)
#pragma unroll
                                for (int i = 0; i < 4; i++) h[t][i] += acc[t][i];
                        }
                        frag_ln(h, w + L.ln2w, w + L.ln2b, p.eps, q, ah, al);
                        float pr[4][4];
#pragma unroll
                        for (int t = 0; t < 4; t++) frag_bias(pr[t], w + L.bpr, t, q);
#pragma unroll 1
                        for (int tf0 = 0; tf0 < 16; tf0 += 4) {
                            float fc[4][4];
#pragma unroll
                            for (int j = 0; j < 4; j++) frag_bias(fc[j], w + L.bfc, tf0 + j, q);
                            gemv_mma<4, 4>(fc, ah, al, w + L.wfc, 128, tf0, g, q);
#pragma unroll
                            for (int j = 0; j < 4; j++) {
#pragma unroll
                                for (int i = 0; i < 4; i++) fc[j][i] = gelu_new_sig(fc[j][i]);
                                uint32_t fh[4], fl[4];
                                frag_c2a(fc[j], fh, fl);
#pragma unroll
                                for (int u = 0; u < 4; u++) {
                                    float w0, w1;
                                    ldb(w + L.wpr, 32, tf0 + j, u, g, q, w0, w1);
                                    mma3(pr[u], fh, fl, w0, w1);
                                }
                            }
                        }
#pragma unroll
                        for (int t = 0; t < 4; t++)
#pragma unroll
                            for (int i = 0; i < 4; i++) h[t][i] += pr[t][i];
#if WS_PROF
                        const long long tp2 = clock64();
                        tA += tp2 - tp1;
#endif
                        if (l == 0) {
                            // ---- step s - 1 is complete: final LayerNorm + linear head, next input ----
                            frag_ln(h, gw, gw + 32, p.eps, q, ah, al);
                            float y[4][4];
#pragma unroll
                            for (int t = 0; t < 4; t++) frag_bias(y[t], gw + 64 + 1024, t, q);
                            gemv_mma<4, 4>(y, ah, al, gw + 64, 32, 0, g, q);
                            const int pnext = min(s, C - 1);
#pragma unroll
                            for (int t = 0; t < 4; t++) {
                                const float2 pe = __ldg(reinterpret_cast<const float2*>(p.pe + pnext * 32 + 8 * t + 2 * q));
#pragma unroll
                                for (int rr = 0; rr < 2; rr++) {
                                    if (br[rr] < p.B)
                                        *reinterpret_cast<float2*>(p.out + (size_t)br[rr] * p.out_stride + (size_t)(s - 1) * 32 +
                                                                   8 * t + 2 * q) = make_float2(y[t][2 * rr], y[t][2 * rr + 1]);
                                    h[t][2 * rr] = y[t][2 * rr] + pe.x;
                                    h[t][2 * rr + 1] = y[t][2 * rr + 1] + pe.y;
                                }
                            }
                        }
                    }
#if WS_PROF
                    const long long tp3 = clock64();
#endif
                    if (n < nsteps) {
                        // ---- first half of layer l of step s: LN1, q / k / v, append to the ring, hand over ----
                        const float* w = wsm + l * L.layer_stride;
                        // LN1(h) goes to the tile as plain FP32 rows: the attention warp computes q, k, v from it
                        frag_ln_store(h, w + L.ln1w, w + L.ln1b, p.eps, q, tile + g * V4QS, tile + (g + 8) * V4QS);
#if WS_PROF
                        const long long tp4 = clock64();
                        tB += tp4 - tp3;
#endif
                        __threadfence_block();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&qready[gi]);
#if WS_PROF
                        tC += clock64() - tp4;
#endif
                    } else if (p.presents != nullptr) {
                        const int Tk = min(p.S, C);
                        const int H = 4, hd = 8;
                        const size_t per_l = (size_t)2 * p.B * H * Tk * hd;
                        const int hh = lane >> 3, d = lane & 7;
                        for (int tg = 0; tg < V4G; tg++) {
                            const int b = grp * V4G + tg;
                            if (b >= p.B) break;
                            for (int rr = 0; rr < NL * 2 * Tk; rr++) {
                                const int t = rr % Tk, lw = rr / Tk;
                                const int which = lw & 1, ll = lw >> 1;
                                const int sl = (p.S - Tk + t) % C;
                                p.presents[(size_t)ll * per_l + ((((size_t)which * p.B + b) * H + hh) * Tk + t) * hd + d] =
                                    kvs[(size_t)tg * kv_per_g + ((size_t)lw * C + sl) * 32 + lane];
                            }
                        }
                        __syncwarp();
                    }
                }
                // the other slot becomes the current one
#pragma unroll
                for (int t = 0; t < 4; t++)
#pragma unroll
                    for (int i = 0; i < 4; i++) { const float x = h[t][i]; h[t][i] = hB[t][i]; hB[t][i] = x; }
                { const uint32_t x = ph; ph = phB; phB = x; }
            }
        }
    }
#if WS_PROF
    if (p.prof != nullptr && blockIdx.x == 0 && lane == 0) {
        p.prof[2 * warp] = twait;
        p.prof[2 * warp + 1] = clock64() - tstart;
        p.prof[32 + 3 * warp] = tA;
        p.prof[32 + 3 * warp + 1] = tB;
        p.prof[32 + 3 * warp + 2] = tC;
    }
#endif
}

// swizzled weight image for the tensor-core kernel: every [K][N] matrix element (k, n) moves to column
// n ^ (((k >> 1) & 3) << 3) of its row; vectors (LN, biases) are copied unchanged.
__global__ void swizzle_weights_kernel(const float* __restrict__ blob, float* __restrict__ out, Gpt2Layout L) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < L.total; idx += gridDim.x * blockDim.x) {
        int base = -1, N = 0, rel = 0;
        if (idx < L.n_layer * L.layer_stride) {
            const int l = idx / L.layer_stride, o = idx - l * L.layer_stride;
            if (o >= L.wqkv && o < L.bqkv) { base = l * L.layer_stride + L.wqkv; N = 96; rel = o - L.wqkv; }
            else if (o >= L.wproj && o < L.bproj) { base = l * L.layer_stride + L.wproj; N = 32; rel = o - L.wproj; }
            else if (o >= L.wfc && o < L.bfc) { base = l * L.layer_stride + L.wfc; N = 128; rel = o - L.wfc; }
            else if (o >= L.wpr && o < L.bpr) { base = l * L.layer_stride + L.wpr; N = 32; rel = o - L.wpr; }
        } else if (idx >= L.wf && idx < L.bf) { base = L.wf; N = 32; rel = idx - L.wf; }
        if (base < 0) { out[idx] = blob[idx]; continue; }
        const int k = rel / N, n = rel - k * N;
        out[base + k * N + (n ^ (((k >> 1) & 3) << 3))] = blob[idx];
    }
}

// ------------------------------------------------------------------------------------------------
// K1 (cluster): small batches of a wide model (Cylinder: E = 128, 6 layers, 4.6 MB of weights).
// The generic kernel makes every CTA stream ALL weights from L2 every step (123 us per step, latency bound).
// Here a thread-block CLUSTER of 8 CTAs owns one group of G trajectories:
//   * CTA rank r owns the output-column slice [r*E/8, (r+1)*E/8) of every GEMV (q, k, v slices alike), so each
//     CTA reads only 1/8 of the weights;
//   * its per-(layer, rank) weight slice is a contiguous image, fetched by TMA bulk copies into a two-slot
//     shared-memory ring (attention half / MLP half of a layer) half a layer AHEAD of use, completion on mbarriers;
//   * after each GEMV the slice results are broadcast to all 8 CTAs with DSMEM stores (mapa + st.shared::cluster)
//     and a cluster barrier (release/acquire) publishes them; every CTA keeps an identical copy of the residual;
//   * (trajectory, head) attention pairs are dealt round-robin to the CTAs; the KV window ring stays in global
//     memory (L2 resident: a few hundred KB).
// ------------------------------------------------------------------------------------------------
constexpr int CLS = 8;                 // CTAs per cluster

struct ClusterLayout {                 // per-(image, rank) float offsets; image i < n_layer: layer i, image n_layer: head
    int E, sl;                         // sl = E / CLS
    // attention half
    int ln1w, ln1b, bqkv, bproj, wqkv, wproj, att_size;
    // MLP half
    int ln2w, ln2b, bfc, bpr, wfc, wpr, mlp_size;
    __host__ __device__ static ClusterLayout make(int E) {
        ClusterLayout L;
        L.E = E; L.sl = E / CLS;
        int o = 0;
        L.ln1w = o; o += E; L.ln1b = o; o += E; L.bqkv = o; o += 3 * L.sl; L.bproj = o; o += L.sl;
        o = (o + 3) & ~3;
        L.wqkv = o; o += E * 3 * L.sl; L.wproj = o; o += E * L.sl;
        L.att_size = (o + 3) & ~3;
        o = 0;
        L.ln2w = o; o += E; L.ln2b = o; o += E; L.bfc = o; o += 4 * L.sl; L.bpr = o; o += L.sl;
        o = (o + 3) & ~3;
        L.wfc = o; o += E * 4 * L.sl; L.wpr = o; o += 4 * E * L.sl;
        L.mlp_size = (o + 3) & ~3;
        return L;
    }
    // the head image reuses the attention-half format: ln1w/ln1b = ln_f, bproj = bf slice, wproj = Wf^T slice
};

__device__ __forceinline__ int cluster_rank() {
    int r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store v at the same shared-memory offset in every CTA of the cluster (including this one)
__device__ __forceinline__ void dsmem_bcast(float* local, float v) {
    const uint32_t l = smem_u32(local);
#pragma unroll
    for (int r = 0; r < CLS; r++) {
        uint32_t ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(l), "r"(r));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
    }
}

// slice GEMV with the weight slice in shared memory: y[g][n] = bias[n] + sum_k aT[k*G+g] * W[k*N + n], n < N (N % 4 == 0).
// A thread owns (k-part, column quad, trajectory): 4 accumulators, one LDS.128 of weights + one LDS of the activation
// per 4 FFMAs, so all NT threads are busy even for the 16-column slices.  K parts are reduced with shuffles when
// several live in one warp, then across warps through `part` ([NT/32][N][G]) by 4 lanes per output.
template <int G, int NT, class Epi>
__device__ __forceinline__ void cta_gemv_smem(const float* W, const float* bias, int K, int N, const float* aT,
                                              float* part, Epi epi) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const int NQ = N >> 2;
    int QL = 1;
    while (QL < NQ) QL <<= 1;
    const int UL = QL * G;                           // lanes per k-slice (power of two)
    int unit, kp, kparts, np, pw;
    if (UL >= 32) {
        const int wpk = UL >> 5;                     // warps per k-slice
        unit = (warp % wpk) * 32 + lane;
        kp = warp / wpk;
        kparts = NW / wpk;
        np = kparts;
        pw = kp;
    } else {
        const int nsub = 32 / UL;
        unit = lane & (UL - 1);
        kp = warp * nsub + lane / UL;
        kparts = NW * nsub;
        np = NW;
        pw = warp;
    }
    const int g = unit % G, nq = unit / G;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (nq < NQ && kp < kparts) {
        const float* wp = W + nq * 4;
        const float* ap = aT + g;
#pragma unroll 4
        for (int k = kp; k < K; k += kparts) {
            const float4 w = *reinterpret_cast<const float4*>(wp + k * N);
            const float a = ap[k * G];
            a0 = fmaf(a, w.x, a0);
            a1 = fmaf(a, w.y, a1);
            a2 = fmaf(a, w.z, a2);
            a3 = fmaf(a, w.w, a3);
        }
    }
    for (int o = UL; o < 32; o <<= 1) {
        a0 += __shfl_xor_sync(FULL, a0, o);
        a1 += __shfl_xor_sync(FULL, a1, o);
        a2 += __shfl_xor_sync(FULL, a2, o);
        a3 += __shfl_xor_sync(FULL, a3, o);
    }
    if (nq < NQ && pw < np && (UL >= 32 || lane < UL)) {
        float* pp = part + ((size_t)pw * N + nq * 4) * G + g;
        pp[0] = a0;
        pp[G] = a1;
        pp[2 * G] = a2;
        pp[3 * G] = a3;
    }
    __syncthreads();
    const int NG = N * G;
    const int sub = tid & 3;
    for (int base = 0; base < NG; base += NT / 4) {   // uniform trip count: the shuffles need whole warps
        const int o = base + (tid >> 2);
        float s2 = 0.f;
        if (o < NG)
            for (int j = sub; j < np; j += 4) s2 += part[(size_t)j * NG + o];
        s2 += __shfl_xor_sync(FULL, s2, 1);
        s2 += __shfl_xor_sync(FULL, s2, 2);
        if (o < NG && sub == 0) {
            const int nn = o / G;
            epi(nn, o - nn * G, s2 + bias[nn]);
        }
    }
}

// attention of one (trajectory, head) pair by one warp, head size 32, window <= 16 slots: every global load of the
// pass is issued up front (two L2 round trips in total), softmax in registers.  q: smem, stride qs; out[d]: lane d.
__device__ __forceinline__ float cluster_attn_hd32(const float* q, int qs, const float* Kr, const float* Vr, int E, int L) {
    const int lane = threadIdx.x & 31;
    const int slot = lane >> 1, half = lane & 1;
    float s = 0.f;
    if (slot < L) {
        const float4* kp4 = reinterpret_cast<const float4*>(Kr + (size_t)slot * E + half * 16);
        const float4 k0 = __ldcg(kp4), k1 = __ldcg(kp4 + 1), k2 = __ldcg(kp4 + 2), k3 = __ldcg(kp4 + 3);
        const float* qq = q + half * 16 * qs;
        s = fmaf(qq[0], k0.x, s); s = fmaf(qq[qs], k0.y, s); s = fmaf(qq[2 * qs], k0.z, s); s = fmaf(qq[3 * qs], k0.w, s);
        s = fmaf(qq[4 * qs], k1.x, s); s = fmaf(qq[5 * qs], k1.y, s); s = fmaf(qq[6 * qs], k1.z, s); s = fmaf(qq[7 * qs], k1.w, s);
        s = fmaf(qq[8 * qs], k2.x, s); s = fmaf(qq[9 * qs], k2.y, s); s = fmaf(qq[10 * qs], k2.z, s); s = fmaf(qq[11 * qs], k2.w, s);
        s = fmaf(qq[12 * qs], k3.x, s); s = fmaf(qq[13 * qs], k3.y, s); s = fmaf(qq[14 * qs], k3.z, s); s = fmaf(qq[15 * qs], k3.w, s);
    }
    // V rows: lane = channel; issue all loads before the softmax math
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = (j < L) ? __ldcg(Vr + (size_t)j * E + lane) : 0.f;
    s += __shfl_xor_sync(FULL, s, 1);
    s = (slot < L) ? s / sqrtf(32.0f) : -INFINITY;
    const float m = warp_max(s);
    float pexp = (slot < L && half == 0) ? expf(s - m) : 0.f;
    const float sum = warp_sum(pexp);
    pexp = pexp / sum;
    float o = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j++) o = fmaf(__shfl_sync(FULL, pexp, 2 * j), v[j], o);
    return o;
}

struct ClusterParams {
    const float* img;          // [n_layer + 1][CLS][att_size + mlp_size]
    const float* pe;
    const float* x0;
    long long x0_stride;
    float* out;
    long long out_stride;
    float* kv;
    float* presents;
    int B, S, use_cache;
    float eps;
    Gpt2Layout lay;
    ClusterLayout cl;
};

template <int G, int NT>
__global__ void __launch_bounds__(NT, 1) rollout_cluster_kernel(ClusterParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t full[2];
    const Gpt2Layout& L = p.lay;
    const ClusterLayout& CLy = p.cl;
    const int E = L.E, H = L.n_head, hd = L.hd, C = L.n_ctx, NL = L.n_layer, sl = CLy.sl;
    const int tid = threadIdx.x, warp = tid >> 5, nw = NT >> 5;
    const int rank = cluster_rank();
    const int grp = blockIdx.x / CLS;
    const int b0 = grp * G;
    float* slotA = sm;                               // attention-half weight slice
    float* slotM = slotA + CLy.att_size;             // MLP-half weight slice
    float* xT = slotM + CLy.mlp_size;                // [E][G] residual (identical in every CTA)
    float* aT = xT + E * G;                          // [E][G] LN output / attention output
    float* qT = aT + E * G;                          // [3E][G]
    float* hT = qT + 3 * E * G;                      // [4E][G]
    float* dT = hT + 4 * E * G;                      // [E][G] broadcast GEMV slice results (residual delta / token)
    float* part = dT + 2 * E * G;                    // [NT/32][4*sl][G]   (dT is followed by the second residual buffer)
    float* pbuf = part + (NT / 32) * 4 * sl * G;     // [nw][C]
    const size_t img_stride = (size_t)CLy.att_size + CLy.mlp_size;
    const size_t kv_per_g = (size_t)NL * 2 * C * E;
    float* kvb = p.kv + (size_t)grp * G * kv_per_g;

    auto img_of = [&](int image) { return p.img + ((size_t)image * CLS + rank) * img_stride; };
    auto fetch = [&](int slot, int image) {           // thread 0 only
        float* dst = slot ? slotM : slotA;
        const float* src = img_of(image) + (slot ? CLy.att_size : 0);
        const uint32_t bytes = (uint32_t)(slot ? CLy.mlp_size : CLy.att_size) * 4u;
        mbar_expect_tx(&full[slot], bytes);
        const uint32_t CH = 32768u;
        for (uint32_t off = 0; off < bytes; off += CH)
            bulk_g2s(reinterpret_cast<char*>(dst) + off, reinterpret_cast<const char*>(src) + off, min(CH, bytes - off),
                     &full[slot]);
    };

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
    }
    __syncthreads();
    if (tid == 0) {
        fetch(0, 0);
        fetch(1, 0);
    }
    uint32_t useA = 0, useM = 0;                      // completed uses of each slot -> wait parity

    for (int i = tid; i < E * G; i += NT) {
        const int c = i / G, g = i - c * G;
        const int b = b0 + g;
        xT[i] = ((b < p.B) ? p.x0[(size_t)b * p.x0_stride + c] : 0.f) + __ldg(p.pe + c);
    }
    cluster_sync_all();                               // all CTAs of the cluster are resident before any DSMEM store

    // Residual update + LayerNorm.  pend: what still has to be applied to the residual from the last exchange
    // (0 nothing, 1 x += dT, 2 x = dT).  Small groups (G <= 2): EVERY warp does the whole update and LayerNorm
    // redundantly — same arithmetic, same values, so the concurrent identical writes are benign — into the other
    // residual buffer (the update is then idempotent) and needs only a __syncwarp: no CTA barrier and no warps idling
    // behind the one that normalises (that wait was 13 % of the step).  Larger groups: one warp per row + barriers.
    float* xN = dT + E * G;                           // second residual buffer (allocated after dT; part follows)
    int pend = 0;
    auto advance_ln = [&](const float* lw, const float* lb) {
        const int lane = tid & 31;
        if constexpr (G <= 2) {
            float* xo = xT;
            float* xn = pend ? xN : xT;
            if (pend)
                for (int i = lane; i < E * G; i += 32) xn[i] = (pend == 1) ? xo[i] + dT[i] : dT[i];
            __syncwarp();
            for (int g = 0; g < G; g++) {
                float s1 = 0.f;
                for (int c = lane; c < E; c += 32) s1 += xn[c * G + g];
                const float mean = warp_sum(s1) / (float)E;
                float v = 0.f;
                for (int c = lane; c < E; c += 32) {
                    const float d = xn[c * G + g] - mean;
                    v = fmaf(d, d, v);
                }
                const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)E + p.eps);
                for (int c = lane; c < E; c += 32) aT[c * G + g] = (xn[c * G + g] - mean) * rstd * lw[c] + lb[c];
            }
            __syncwarp();
            if (pend) {                               // swap the residual buffers (uniform across the CTA)
                xN = xo;
                xT = xn;
            }
        } else {
            if (pend) {
                for (int i = tid; i < E * G; i += NT) xT[i] = (pend == 1) ? xT[i] + dT[i] : dT[i];
                __syncthreads();
            }
            for (int g = warp; g < G; g += nw) {
                float s1 = 0.f;
                for (int c = lane; c < E; c += 32) s1 += xT[c * G + g];
                const float mean = warp_sum(s1) / (float)E;
                float v = 0.f;
                for (int c = lane; c < E; c += 32) {
                    const float d = xT[c * G + g] - mean;
                    v = fmaf(d, d, v);
                }
                const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)E + p.eps);
                for (int c = lane; c < E; c += 32) aT[c * G + g] = (xT[c * G + g] - mean) * rstd * lw[c] + lb[c];
            }
            __syncthreads();
        }
        pend = 0;
    };

    for (int s = 0; s < p.S; s++) {
        const int Lk = p.use_cache ? min(s + 1, C) : 1;
        const int slot = p.use_cache ? (s % C) : 0;
        for (int l = 0; l <= NL; l++) {
            // ================= attention half (or the final head when l == NL) =================
            mbar_wait(&full[0], useA & 1);
            useA++;
            advance_ln(slotA + CLy.ln1w, slotA + CLy.ln1b);
            if (l == NL) {
                // x' = LN_f(x) Wf^T + bf : this CTA's slice -> every CTA, and to the output
                const int pnext = p.use_cache ? min(s + 1, C - 1) : 0;
                cta_gemv_smem<G, NT>(slotA + CLy.wproj, slotA + CLy.bproj, E, sl, aT, part, [&](int n, int g, float v) {
                    const int col = rank * sl + n, b = b0 + g;
                    if (b < p.B) p.out[(size_t)b * p.out_stride + (size_t)s * E + col] = v;
                    dsmem_bcast(dT + col * G + g, v + __ldg(p.pe + pnext * E + col));
                });
                cluster_sync_all();
                pend = 2;                             // next step starts from x = dT
                if (tid == 0 && s + 1 < p.S) fetch(0, 0);
                break;
            }
            cta_gemv_smem<G, NT>(slotA + CLy.wqkv, slotA + CLy.bqkv, E, 3 * sl, aT, part, [&](int n, int g, float v) {
                const int which = n / sl, col = rank * sl + (n - which * sl);       // q | k | v slices
                dsmem_bcast(qT + (which * E + col) * G + g, v);
                if (which) kvb[(size_t)g * kv_per_g + ((size_t)(l * 2 + which - 1) * C + slot) * E + col] = v;
            });
            cluster_sync_all();
            // (trajectory, head) pairs dealt round-robin: pair p -> CTA p % CLS, warp (p / CLS) % nw
            for (int pair = rank + CLS * warp; pair < G * H; pair += CLS * nw) {
                const int g = pair / H, hh = pair - g * H;
                const float* Kr = kvb + (size_t)g * kv_per_g + (size_t)(l * 2 + 0) * C * E + hh * hd;
                const float* Vr = kvb + (size_t)g * kv_per_g + (size_t)(l * 2 + 1) * C * E + hh * hd;
                if (hd == 32 && C <= 16) {
                    const float o = cluster_attn_hd32(qT + (hh * hd) * G + g, G, Kr, Vr, E, Lk);
                    dsmem_bcast(aT + (hh * hd + (tid & 31)) * G + g, o);
                } else {
                    float* ob = hT + (hh * hd) * G + g;     // local staging, then broadcast
                    warp_attention(qT + (hh * hd) * G + g, G, Kr, Vr, E, hd, Lk, pbuf + warp * C, ob, G, nullptr);
                    for (int d = (tid & 31); d < hd; d += 32) dsmem_bcast(aT + (hh * hd + d) * G + g, ob[d * G]);
                }
            }
            cluster_sync_all();
            cta_gemv_smem<G, NT>(slotA + CLy.wproj, slotA + CLy.bproj, E, sl, aT, part,
                             [&](int n, int g, float v) { dsmem_bcast(dT + (rank * sl + n) * G + g, v); });
            cluster_sync_all();                       // also: every thread of this CTA is done with slot A
            if (tid == 0) fetch(0, (l + 1 <= NL) ? l + 1 : 0);
            pend = 1;
            // ================= MLP half =================
            mbar_wait(&full[1], useM & 1);
            useM++;
            advance_ln(slotM + CLy.ln2w, slotM + CLy.ln2b);
            cta_gemv_smem<G, NT>(slotM + CLy.wfc, slotM + CLy.bfc, E, 4 * sl, aT, part, [&](int n, int g, float v) {
                dsmem_bcast(hT + (rank * 4 * sl + n) * G + g, gelu_new(v));
            });
            cluster_sync_all();
            cta_gemv_smem<G, NT>(slotM + CLy.wpr, slotM + CLy.bpr, 4 * E, sl, hT, part,
                             [&](int n, int g, float v) { dsmem_bcast(dT + (rank * sl + n) * G + g, v); });
            cluster_sync_all();
            if (tid == 0 && !(s + 1 == p.S && l + 1 == NL)) fetch(1, (l + 1 < NL) ? l + 1 : 0);
            pend = 1;
        }
    }
    if (p.presents != nullptr && p.use_cache && rank == 0) {
        const int Tk = min(p.S, C);
        const size_t per_l = (size_t)2 * p.B * H * Tk * hd;
        for (int g = 0; g < G; g++) {
            const int b = b0 + g;
            if (b >= p.B) break;
            for (int i = tid; i < NL * 2 * Tk * E; i += NT) {
                const int e = i % E;
                int r = i / E;
                const int t = r % Tk;
                r /= Tk;
                const int which = r & 1, l = r >> 1;
                const int sl2 = (p.S - Tk + t) % C;
                const int hh = e / hd, d = e - hh * hd;
                p.presents[(size_t)l * per_l + ((((size_t)which * p.B + b) * H + hh) * Tk + t) * hd + d] =
                    kvb[(size_t)g * kv_per_g + ((size_t)(l * 2 + which) * C + sl2) * E + e];
            }
        }
    }
    cluster_sync_all();                               // no CTA may exit while peers can still store into its smem
}

// ------------------------------------------------------------------------------------------------
// K1 (cluster, asynchronous exchange): the same decomposition as rollout_cluster_kernel for groups of 1-2
// trajectories, but no cluster-wide barrier inside the step.  Measured on the barrier version: a step is 31 exchange
// phases of ~2 us, of which ~0.5 us is barrier.cluster and ~0.3 us the release fence waiting for the remote stores.
// Here every exchanged value travels as `st.async ... mbarrier::complete_tx::bytes` (a DSMEM store that credits its
// 4 bytes to an mbarrier of the DESTINATION CTA); each CTA arms one of three rotating mbarriers with the byte count
// the phase owes it and waits on that: point-to-point, no fence, a CTA proceeds as soon as ITS inputs are complete.
// Buffers are reused only after an all-to-all phase in between (see the phase list in the body), the (trajectory,
// head) owner CTA appends the k/v rows to the window itself (so the ring is written and read by one SM), and the
// small groups use the per-warp redundant LayerNorm (no CTA barrier outside the GEMV reduction).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dsmem_bcast_async(float* local, float v, uint64_t* bar) {
    const uint32_t l = smem_u32(local), b = smem_u32(bar);
#pragma unroll
    for (int r = 0; r < CLS; r++) {
        uint32_t ra, rb;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(l), "r"(r));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(b), "r"(r));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra),
                     "r"(__float_as_uint(v)), "r"(rb)
                     : "memory");
    }
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TRPX_WAITC_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TRPX_DONEC_%=;\n"
        "bra TRPX_WAITC_%=;\n"
        "TRPX_DONEC_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <int G, int NT>
__global__ void __launch_bounds__(NT, NT <= 256 ? 2 : 1) rollout_cluster_async_kernel(ClusterParams p) {
    static_assert(G <= 2, "the asynchronous exchange kernel uses the per-warp redundant LayerNorm (small groups)");
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t full[2];         // weight ring
    __shared__ __align__(8) uint64_t xb[3];           // exchange phases, rotating: phase ph uses xb[ph % 3]
    const Gpt2Layout& L = p.lay;
    const ClusterLayout& CLy = p.cl;
    const int E = L.E, H = L.n_head, hd = L.hd, C = L.n_ctx, NL = L.n_layer, sl = CLy.sl;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
    const int rank = cluster_rank();
    const int grp = blockIdx.x / CLS;
    const int b0 = grp * G;
    const int EG = E * G;
    float* slotA = sm;
    float* slotM = slotA + CLy.att_size;
    float* xT = slotM + CLy.mlp_size;                // residual, two buffers
    float* xN = xT + EG;
    float* aT = xN + EG;                             // LN output (local)
    float* qT = aT + EG;                             // [3E][G]  phase 1 (all -> all)
    float* oT = qT + 3 * EG;                         // [E][G]   phase 2 (owners -> all)
    float* dT = oT + EG;                             // [E][G]   phases 3 and 5
    float* hT = dT + EG;                             // [4E][G]  phase 4
    float* yT = hT + 4 * EG;                         // [E][G]   head
    float* part = yT + EG;                           // [NT/32][4*sl][G]
    float* pbuf = part + (NT / 32) * 4 * sl * G;     // [nw][C]
    const size_t img_stride = (size_t)CLy.att_size + CLy.mlp_size;
    const size_t kv_per_g = (size_t)NL * 2 * C * E;
    float* kvb = p.kv + (size_t)grp * G * kv_per_g;

    auto img_of = [&](int image) { return p.img + ((size_t)image * CLS + rank) * img_stride; };
    auto fetch = [&](int slot, int image) {           // thread 0 only
        float* dst = slot ? slotM : slotA;
        const float* src = img_of(image) + (slot ? CLy.att_size : 0);
        const uint32_t bytes = (uint32_t)(slot ? CLy.mlp_size : CLy.att_size) * 4u;
        mbar_expect_tx(&full[slot], bytes);
        const uint32_t CH = 32768u;
        for (uint32_t off = 0; off < bytes; off += CH)
            bulk_g2s(reinterpret_cast<char*>(dst) + off, reinterpret_cast<const char*>(src) + off, min(CH, bytes - off),
                     &full[slot]);
    };
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&xb[0], 1);
        mbar_init(&xb[1], 1);
        mbar_init(&xb[2], 1);
    }
    __syncthreads();
    if (tid == 0) {
        fetch(0, 0);
        fetch(1, 0);
    }
    uint32_t useA = 0, useM = 0;
    // Exchange phase counter: barrier xb[ph % 3], parity (ph / 3) & 1.  Three barriers, not two: phase 2 is sent by
    // the pair owners only, so a CTA that owns no pair could otherwise receive phase-3 bytes on the barrier it is
    // still waiting on for phase 1; with a rotation of three a sender of phase ph+3 has always (transitively) seen
    // this CTA's own send of phase ph+1 or ph+2, which it issues after its wait of phase ph.
    uint32_t ph = 0;
    for (int i = tid; i < EG; i += NT) {
        const int c = i / G, g = i - c * G;
        const int b = b0 + g;
        xT[i] = ((b < p.B) ? p.x0[(size_t)b * p.x0_stride + c] : 0.f) + __ldg(p.pe + c);
    }
    cluster_sync_all();                               // barriers initialised and CTAs resident before any remote store

    // arm the barrier of the phase that starts now with the bytes this CTA will receive in it
    auto arm = [&](uint32_t bytes) {
        if (tid == 0) mbar_expect_tx(&xb[ph % 3], bytes);
    };
    auto wait_phase = [&]() {
        mbar_wait_cluster(&xb[ph % 3], (ph / 3) & 1);
        ph++;
    };
    int pend = 0;                                     // 0 nothing, 1 x += dT, 2 x = yT
    auto advance_ln = [&](const float* lw, const float* lb) {
        float* xo = xT;
        float* xn = pend ? xN : xT;
        if (pend)
            for (int i = lane; i < EG; i += 32) xn[i] = (pend == 1) ? xo[i] + dT[i] : yT[i];
        __syncwarp();
        for (int g = 0; g < G; g++) {
            float s1 = 0.f;
            for (int c = lane; c < E; c += 32) s1 += xn[c * G + g];
            const float mean = warp_sum(s1) / (float)E;
            float v = 0.f;
            for (int c = lane; c < E; c += 32) {
                const float d = xn[c * G + g] - mean;
                v = fmaf(d, d, v);
            }
            const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)E + p.eps);
            for (int c = lane; c < E; c += 32) aT[c * G + g] = (xn[c * G + g] - mean) * rstd * lw[c] + lb[c];
        }
        __syncwarp();
        if (pend) {
            xN = xo;
            xT = xn;
        }
        pend = 0;
    };

    for (int s = 0; s < p.S; s++) {
        const int Lk = p.use_cache ? min(s + 1, C) : 1;
        const int slot = p.use_cache ? (s % C) : 0;
        for (int l = 0; l <= NL; l++) {
            mbar_wait(&full[0], useA & 1);
            useA++;
            if (l == NL) {
                // ---- head: x' = LN_f(x) Wf^T + bf ----
                arm((uint32_t)EG * 4u);
                advance_ln(slotA + CLy.ln1w, slotA + CLy.ln1b);
                const int pnext = p.use_cache ? min(s + 1, C - 1) : 0;
                uint64_t* bar = &xb[ph % 3];
                cta_gemv_smem<G, NT>(slotA + CLy.wproj, slotA + CLy.bproj, E, sl, aT, part, [&](int n, int g, float v) {
                    const int col = rank * sl + n, b = b0 + g;
                    if (b < p.B) p.out[(size_t)b * p.out_stride + (size_t)s * E + col] = v;
                    dsmem_bcast_async(yT + col * G + g, v + __ldg(p.pe + pnext * E + col), bar);
                });
                wait_phase();
                pend = 2;
                __syncthreads();                      // every warp is done with slot A
                if (tid == 0 && s + 1 < p.S) fetch(0, 0);
                break;
            }
            // ---- phase 1: q | k | v slices, all -> all ----
            arm((uint32_t)3 * EG * 4u);
            advance_ln(slotA + CLy.ln1w, slotA + CLy.ln1b);
            {
                uint64_t* bar = &xb[ph % 3];
                cta_gemv_smem<G, NT>(slotA + CLy.wqkv, slotA + CLy.bqkv, E, 3 * sl, aT, part, [&](int n, int g, float v) {
                    const int which = n / sl, col = rank * sl + (n - which * sl);
                    dsmem_bcast_async(qT + (which * E + col) * G + g, v, bar);
                });
            }
            wait_phase();
            // ---- phase 2: attention outputs, owners -> all ----
            arm((uint32_t)EG * 4u);
            {
                uint64_t* bar = &xb[ph % 3];
                for (int pair = rank + CLS * warp; pair < G * H; pair += CLS * nw) {
                    const int g = pair / H, hh = pair - g * H;
                    float* Kr = kvb + (size_t)g * kv_per_g + (size_t)(l * 2 + 0) * C * E + hh * hd;
                    float* Vr = kvb + (size_t)g * kv_per_g + (size_t)(l * 2 + 1) * C * E + hh * hd;
                    for (int d = lane; d < hd; d += 32) {   // the owner appends this pair's k / v row to the window
                        Kr[(size_t)slot * E + d] = qT[(E + hh * hd + d) * G + g];
                        Vr[(size_t)slot * E + d] = qT[(2 * E + hh * hd + d) * G + g];
                    }
                    __syncwarp();
                    if (hd == 32 && C <= 16) {
                        const float o = cluster_attn_hd32(qT + (hh * hd) * G + g, G, Kr, Vr, E, Lk);
                        dsmem_bcast_async(oT + (hh * hd + lane) * G + g, o, bar);
                    } else {
                        float* ob = aT + (hh * hd) * G + g;   // LN output is consumed: local staging, then broadcast
                        warp_attention(qT + (hh * hd) * G + g, G, Kr, Vr, E, hd, Lk, pbuf + warp * C, ob, G, nullptr);
                        for (int d = lane; d < hd; d += 32) dsmem_bcast_async(oT + (hh * hd + d) * G + g, ob[d * G], bar);
                    }
                }
            }
            wait_phase();
            // ---- phase 3: projection slices ----
            arm((uint32_t)EG * 4u);
            {
                uint64_t* bar = &xb[ph % 3];
                cta_gemv_smem<G, NT>(slotA + CLy.wproj, slotA + CLy.bproj, E, sl, oT, part,
                                     [&](int n, int g, float v) { dsmem_bcast_async(dT + (rank * sl + n) * G + g, v, bar); });
            }
            wait_phase();
            __syncthreads();                          // every warp is done with slot A
            if (tid == 0) fetch(0, l + 1);
            pend = 1;
            // ---- phase 4: MLP hidden slices ----
            mbar_wait(&full[1], useM & 1);
            useM++;
            arm((uint32_t)4 * EG * 4u);
            advance_ln(slotM + CLy.ln2w, slotM + CLy.ln2b);
            {
                uint64_t* bar = &xb[ph % 3];
                cta_gemv_smem<G, NT>(slotM + CLy.wfc, slotM + CLy.bfc, E, 4 * sl, aT, part, [&](int n, int g, float v) {
                    dsmem_bcast_async(hT + (rank * 4 * sl + n) * G + g, gelu_new(v), bar);
                });
            }
            wait_phase();
            // ---- phase 5: MLP output slices ----
            arm((uint32_t)EG * 4u);
            {
                uint64_t* bar = &xb[ph % 3];
                cta_gemv_smem<G, NT>(slotM + CLy.wpr, slotM + CLy.bpr, 4 * E, sl, hT, part,
                                     [&](int n, int g, float v) { dsmem_bcast_async(dT + (rank * sl + n) * G + g, v, bar); });
            }
            wait_phase();
            __syncthreads();                          // every warp is done with slot M
            if (tid == 0 && !(s + 1 == p.S && l + 1 == NL)) fetch(1, (l + 1 < NL) ? l + 1 : 0);
            pend = 1;
        }
    }
    cluster_sync_all();                               // ring rows of every owner visible; nobody stores into a peer any more
    if (p.presents != nullptr && p.use_cache && rank == 0) {
        const int Tk = min(p.S, C);
        const size_t per_l = (size_t)2 * p.B * H * Tk * hd;
        for (int g = 0; g < G; g++) {
            const int b = b0 + g;
            if (b >= p.B) break;
            for (int i = tid; i < NL * 2 * Tk * E; i += NT) {
                const int e = i % E;
                int r = i / E;
                const int t = r % Tk;
                r /= Tk;
                const int which = r & 1, l = r >> 1;
                const int sl2 = (p.S - Tk + t) % C;
                const int hh = e / hd, d = e - hh * hd;
                p.presents[(size_t)l * per_l + ((((size_t)which * p.B + b) * H + hh) * Tk + t) * hd + d] =
                    __ldcg(kvb + (size_t)g * kv_per_g + ((size_t)(l * 2 + which) * C + sl2) * E + e);
            }
        }
    }
}

// per-(image, rank) weight images for the cluster kernel
__global__ void cluster_image_kernel(const float* __restrict__ blob, float* __restrict__ img, Gpt2Layout L, ClusterLayout C) {
    const int E = L.E, sl = C.sl;
    const size_t stride = (size_t)C.att_size + C.mlp_size;
    const size_t total = (size_t)(L.n_layer + 1) * CLS * stride;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(i % stride);
        const int rank = (int)((i / stride) % CLS), image = (int)(i / (stride * CLS));
        float v = 0.f;
        if (image < L.n_layer) {
            const float* w = blob + (size_t)image * L.layer_stride;
            if (o < C.att_size) {
                if (o < C.ln1b) v = w[L.ln1w + o];
                else if (o < C.bqkv) v = w[L.ln1b + o - C.ln1b];
                else if (o < C.bproj) { const int n = o - C.bqkv, wh = n / sl; v = w[L.bqkv + wh * E + rank * sl + n % sl]; }
                else if (o < C.bproj + sl) v = w[L.bproj + rank * sl + o - C.bproj];
                else if (o >= C.wqkv && o < C.wproj) {
                    const int r = o - C.wqkv, k = r / (3 * sl), n = r % (3 * sl), wh = n / sl;
                    v = w[L.wqkv + k * 3 * E + wh * E + rank * sl + n % sl];
                } else if (o >= C.wproj && o < C.wproj + E * sl) {
                    const int r = o - C.wproj, k = r / sl, n = r % sl;
                    v = w[L.wproj + k * E + rank * sl + n];
                }
            } else {
                const int m = o - C.att_size;
                if (m < C.ln2b) v = w[L.ln2w + m];
                else if (m < C.bfc) v = w[L.ln2b + m - C.ln2b];
                else if (m < C.bpr) v = w[L.bfc + rank * 4 * sl + m - C.bfc];
                else if (m < C.bpr + sl) v = w[L.bpr + rank * sl + m - C.bpr];
                else if (m >= C.wfc && m < C.wpr) {
                    const int r = m - C.wfc, k = r / (4 * sl), n = r % (4 * sl);
                    v = w[L.wfc + k * 4 * E + rank * 4 * sl + n];
                } else if (m >= C.wpr && m < C.wpr + 4 * E * sl) {
                    const int r = m - C.wpr, k = r / sl, n = r % sl;
                    v = w[L.wpr + k * E + rank * sl + n];
                }
            }
        } else if (o < C.att_size) {                   // head image in the attention-half format
            if (o < C.ln1b) v = blob[L.lnfw + o];
            else if (o < C.bqkv) v = blob[L.lnfb + o - C.ln1b];
            else if (o >= C.bproj && o < C.bproj + sl) v = blob[L.bf + rank * sl + o - C.bproj];
            else if (o >= C.wproj && o < C.wproj + E * sl) {
                const int r = o - C.wproj, k = r / sl, n = r % sl;
                v = blob[L.wf + k * E + rank * sl + n];
            }
        }
        img[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: teacher-forced forward of T tokens (optional past of Tp keys), one CTA per batch element.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_FWD_LAYERS = 32;

struct ForwardParams {
    const float* blob;
    const float* pe;
    const float* x;
    const float* past[MAX_FWD_LAYERS];      // per-layer pointers (passed by value), valid if has_past
    float* hidden;
    float* presents[MAX_FWD_LAYERS];
    float* attn[MAX_FWD_LAYERS];
    int has_past, has_presents, has_attn;
    int B, T, Tp;
    float eps;
    Gpt2Layout lay;
};

constexpr int TG = 8;              // tokens per GEMV register block

__global__ void __launch_bounds__(GEN_THREADS) forward_kernel(ForwardParams p) {
    extern __shared__ __align__(16) float sm[];
    const Gpt2Layout& L = p.lay;
    const int E = L.E, H = L.n_head, hd = L.hd, C = L.n_ctx;
    const int T = p.T, Tp = p.Tp, Tk = Tp + T;
    const int nch = (T + TG - 1) / TG, TP = nch * TG;
    float* x = sm;                         // [TP][E]   residual, row-major
    float* a = x + TP * E;                 // [nch][E][TG]  gemv input (chunked, transposed)
    float* q = a + TP * E;                 // [TP][E]   queries, row-major
    float* ks = q + TP * E;                // [C][E]
    float* vs = ks + C * E;                // [C][E]
    float* hb = vs + C * E;                // [nch][4E][TG]
    float* part = hb + TP * 4 * E;         // [blockDim*TG]
    float* pbuf = part + GEN_THREADS * TG; // [nwarps][C]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    const int b = blockIdx.x;

    for (int i = tid; i < TP * E; i += blockDim.x) {
        const int t = i / E, c = i - t * E;
        // position id = past_length + t as float (phys_transformer_gpt2.py:198)
        x[i] = (t < T) ? p.x[((size_t)b * T + t) * E + c] + __ldg(p.pe + (Tp + t) * E + c) : 0.f;
    }
    __syncthreads();

    auto ln_to_a = [&](const float* lw, const float* lb) {
        for (int t = warp; t < TP; t += nw) {
            float s = 0.f;
            for (int c = lane; c < E; c += 32) s += x[t * E + c];
            const float mean = warp_sum(s) / (float)E;
            float v = 0.f;
            for (int c = lane; c < E; c += 32) {
                const float d = x[t * E + c] - mean;
                v = fmaf(d, d, v);
            }
            const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)E + p.eps);
            float* dst = a + (t / TG) * E * TG + (t % TG);
            for (int c = lane; c < E; c += 32)
                dst[c * TG] = (x[t * E + c] - mean) * rstd * __ldg(lw + c) + __ldg(lb + c);
        }
    };

    for (int l = 0; l < L.n_layer; l++) {
        const float* w = p.blob + (size_t)l * L.layer_stride;
        // past keys/values -> window rows [0, Tp)
        if (Tp > 0) {
            const float* pk = p.past[l];       // [2][B][H][Tp][hd]
            for (int i = tid; i < 2 * Tp * E; i += blockDim.x) {
                const int which = i / (Tp * E);
                int r = i - which * Tp * E;
                const int t = r / E, e = r - t * E;
                const int hh = e / hd, d = e - hh * hd;
                const float v = pk[((((size_t)which * p.B + b) * H + hh) * Tp + t) * hd + d];
                (which ? vs : ks)[t * E + e] = v;
            }
        }
        ln_to_a(w + L.ln1w, w + L.ln1b);
        __syncthreads();
        for (int ch = 0; ch < nch; ch++) {
            block_gemv<TG>(w + L.wqkv, w + L.bqkv, E, 3 * E, a + ch * E * TG, part, [&](int n, int g, float v) {
                const int t = ch * TG + g;
                if (t >= T) return;
                if (n < E) q[t * E + n] = v;
                else if (n < 2 * E) ks[(Tp + t) * E + (n - E)] = v;
                else vs[(Tp + t) * E + (n - 2 * E)] = v;
            });
            __syncthreads();
        }
        if (p.has_presents) {
            float* pr = p.presents[l];         // [2][B][H][Tk][hd]
            for (int i = tid; i < 2 * Tk * E; i += blockDim.x) {
                const int which = i / (Tk * E);
                int r = i - which * Tk * E;
                const int t = r / E, e = r - t * E;
                const int hh = e / hd, d = e - hh * hd;
                pr[((((size_t)which * p.B + b) * H + hh) * Tk + t) * hd + d] = (which ? vs : ks)[t * E + e];
            }
        }
        // causal attention: query t sees keys [0, Tp + t]
        for (int pair = warp; pair < T * H; pair += nw) {
            const int t = pair / H, hh = pair - t * H;
            float* ao = nullptr;
            if (p.has_attn) {
                ao = p.attn[l] + (((size_t)b * H + hh) * T + t) * Tk;
                for (int j = Tp + t + 1 + lane; j < Tk; j += 32) ao[j] = 0.f;   // masked: exp(-1e4 - max) == 0
            }
            warp_attention(q + t * E + hh * hd, 1, ks + hh * hd, vs + hh * hd, E, hd, Tp + t + 1, pbuf + warp * C,
                           a + (t / TG) * E * TG + (hh * hd) * TG + (t % TG), TG, ao);
        }
        __syncthreads();
        for (int ch = 0; ch < nch; ch++) {
            block_gemv<TG>(w + L.wproj, w + L.bproj, E, E, a + ch * E * TG, part, [&](int n, int g, float v) {
                const int t = ch * TG + g;
                if (t < T) x[t * E + n] += v;
            });
            __syncthreads();
        }
        ln_to_a(w + L.ln2w, w + L.ln2b);
        __syncthreads();
        for (int ch = 0; ch < nch; ch++) {
            block_gemv<TG>(w + L.wfc, w + L.bfc, E, 4 * E, a + ch * E * TG, part, [&](int n, int g, float v) {
                hb[ch * 4 * E * TG + n * TG + g] = gelu_new(v);
            });
            __syncthreads();
        }
        for (int ch = 0; ch < nch; ch++) {
            block_gemv<TG>(w + L.wpr, w + L.bpr, 4 * E, E, hb + ch * 4 * E * TG, part, [&](int n, int g, float v) {
                const int t = ch * TG + g;
                if (t < T) x[t * E + n] += v;
            });
            __syncthreads();
        }
    }
    ln_to_a(p.blob + L.lnfw, p.blob + L.lnfb);
    __syncthreads();
    for (int ch = 0; ch < nch; ch++) {
        block_gemv<TG>(p.blob + L.wf, p.blob + L.bf, E, E, a + ch * E * TG, part, [&](int n, int g, float v) {
            const int t = ch * TG + g;
            if (t < T) p.hidden[((size_t)b * T + t) * E + n] = v;
        });
        __syncthreads();
    }
}

// ---- teacher-forced forward on tensor cores (E = 32, 4 heads of 8, T + Tp <= 64) ---------------------------------
// forward_kernel above runs a sequence through block-wide GEMVs of 8 tokens with a CTA barrier after each (160 of
// them for T = 64): 3 TFLOP/s.  Here a sequence is four 16-token M tiles, one per warp; every dense op of a layer is a
// warp-level 3xTF32 product (mma.sync.m16n8k8, a_lo*w_hi + a_hi*w_lo + a_hi*w_hi, FP32 accumulate: ~2^-21 relative, the
// FP32 parity bar holds) on the warp's own rows — activations in shared memory rows that only this warp touches,
// weights read through L1 — so the only CTA barriers are the two around the attention (keys / values of all tokens).
constexpr int FT_ROWS = 64, FT_LDX = 36, FT_LDH = 132, FT_LDK = 33, FT_THREADS = 128;

__device__ __forceinline__ void ft_split(float x, float& hi, float& lo) {
    hi = __uint_as_float(tf32_rna_bits(x));
    lo = x - hi;
}
__device__ __forceinline__ void ft_mma(float (&c)[4], const float (&a)[4], float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// C[16 x N] = A[16 x K] (shared memory, row stride lda) . W[K x N] (global, row-major) + bias; epi(row, col, value)
template <int K, int N, class Epi>
__device__ __forceinline__ void ft_warp_gemm(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                             const float* __restrict__ bias, int lane, Epi epi) {
    const int g = lane >> 2, q = lane & 3;
    float acc[N / 8][4];
#pragma unroll
    for (int nt = 0; nt < N / 8; nt++) {
        const float b0 = __ldg(bias + nt * 8 + 2 * q), b1 = __ldg(bias + nt * 8 + 2 * q + 1);
        acc[nt][0] = b0; acc[nt][1] = b1; acc[nt][2] = b0; acc[nt][3] = b1;
    }
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 8) {
        float ah[4], al[4];
        ft_split(A[g * lda + k0 + q], ah[0], al[0]);
        ft_split(A[(g + 8) * lda + k0 + q], ah[1], al[1]);
        ft_split(A[g * lda + k0 + q + 4], ah[2], al[2]);
        ft_split(A[(g + 8) * lda + k0 + q + 4], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < N / 8; nt++) {
            float bh0, bl0, bh1, bl1;
            ft_split(__ldg(W + (size_t)(k0 + q) * N + nt * 8 + g), bh0, bl0);
            ft_split(__ldg(W + (size_t)(k0 + q + 4) * N + nt * 8 + g), bh1, bl1);
            ft_mma(acc[nt], al, bh0, bh1);
            ft_mma(acc[nt], ah, bl0, bl1);
            ft_mma(acc[nt], ah, bh0, bh1);
        }
    }
#pragma unroll
    for (int nt = 0; nt < N / 8; nt++) {
        epi(g, nt * 8 + 2 * q, acc[nt][0]);
        epi(g, nt * 8 + 2 * q + 1, acc[nt][1]);
        epi(g + 8, nt * 8 + 2 * q, acc[nt][2]);
        epi(g + 8, nt * 8 + 2 * q + 1, acc[nt][3]);
    }
}

// LayerNorm of the warp's 16 rows of x (stride FT_LDX) into dst (stride ldd): two lanes per row, 16 channels each
__device__ __forceinline__ void ft_warp_ln(const float* __restrict__ x, float* __restrict__ dst, int ldd,
                                           const float* __restrict__ lw, const float* __restrict__ lb, float eps, int lane) {
    const int r = lane >> 1, c0 = (lane & 1) * 16;
    float v[16], s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) { v[i] = x[r * FT_LDX + c0 + i]; s += v[i]; }
    s += __shfl_xor_sync(FULL, s, 1);
    const float mean = s * (1.0f / 32.0f);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) { const float d = v[i] - mean; var = fmaf(d, d, var); }
    var += __shfl_xor_sync(FULL, var, 1);
    const float rstd = 1.0f / sqrtf(var * (1.0f / 32.0f) + eps);
#pragma unroll
    for (int i = 0; i < 16; i++) dst[r * ldd + c0 + i] = (v[i] - mean) * rstd * __ldg(lw + c0 + i) + __ldg(lb + c0 + i);
}

__global__ void __launch_bounds__(FT_THREADS, 4) forward_e32_tc_kernel(ForwardParams p) {
    extern __shared__ __align__(16) float sm[];
    const Gpt2Layout& L = p.lay;
    constexpr int E = 32, H = 4, HD = 8;
    const int T = p.T, Tp = p.Tp, Tk = Tp + T;
    float* x = sm;                              // [64][36]  residual stream
    float* a = x + FT_ROWS * FT_LDX;            // [64][36]  GEMM input (LayerNorm output, attention output)
    float* hb = a + FT_ROWS * FT_LDX;           // [64][132] MLP hidden; its first rows are ALSO the q / k / v tiles of the
    float* qs = hb;                             // [64][33]  attention phase (the two never overlap in time: a CTA barrier
    float* ks = qs + FT_ROWS * FT_LDK;          // [64][33]  separates the attention from the MLP and the MLP from the next
    float* vs = ks + FT_ROWS * FT_LDK;          // [64][33]  layer's q|k|v); keys of the window: past rows first
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    const int r0 = warp * 16;                   // this warp's token rows
    float* xw = x + r0 * FT_LDX;
    float* aw = a + r0 * FT_LDX;
    float* hw = hb + r0 * FT_LDH;

    for (int i = tid; i < FT_ROWS * E; i += FT_THREADS) {
        const int t = i >> 5, c = i & 31;
        x[t * FT_LDX + c] = (t < T) ? p.x[((size_t)b * T + t) * E + c] + __ldg(p.pe + (Tp + t) * E + c) : 0.f;
    }
    __syncthreads();
    for (int l = 0; l < L.n_layer; l++) {
        const float* w = p.blob + (size_t)l * L.layer_stride;
        if (Tp > 0) {
            if (l > 0) __syncthreads();         // (the past rows land in the tile the previous MLP used)
            const float* pk = p.past[l];       // [2][B][H][Tp][hd]
            for (int i = tid; i < 2 * Tp * E; i += FT_THREADS) {
                const int which = i / (Tp * E);
                int r = i - which * Tp * E;
                const int t = r >> 5, e = r & 31;
                const int hh = e >> 3, d = e & 7;
                (which ? vs : ks)[t * FT_LDK + e] = pk[((((size_t)which * p.B + b) * H + hh) * Tp + t) * HD + d];
            }
        }
        ft_warp_ln(xw, aw, FT_LDX, w + L.ln1w, w + L.ln1b, p.eps, lane);
        __syncwarp();
        if (l > 0) __syncthreads();             // every warp is done with the previous layer's MLP hidden tile (aliases q|k|v)
        ft_warp_gemm<E, 3 * E>(aw, FT_LDX, w + L.wqkv, w + L.bqkv, lane, [&](int r, int n, float v) {
            const int t = r0 + r;
            if (t >= T) return;
            if (n < E) qs[t * FT_LDK + n] = v;
            else if (n < 2 * E) ks[(Tp + t) * FT_LDK + (n - E)] = v;
            else vs[(Tp + t) * FT_LDK + (n - 2 * E)] = v;
        });
        __syncthreads();                        // keys / values of every token are in place
        if (p.has_presents) {
            float* pr = p.presents[l];         // [2][B][H][Tk][hd]
            for (int i = tid; i < 2 * Tk * E; i += FT_THREADS) {
                const int which = i / (Tk * E);
                int r = i - which * Tk * E;
                const int t = r >> 5, e = r & 31;
                const int hh = e >> 3, d = e & 7;
                pr[((((size_t)which * p.B + b) * H + hh) * Tk + t) * HD + d] = (which ? vs : ks)[t * FT_LDK + e];
            }
        }
        // causal attention: query t sees keys [0, Tp + t].  A thread owns one head of the token pair (i, T - 1 - i): every
        // thread of the CTA walks T + 1 (+ 2 Tp) keys, whatever its tokens (a thread per token walked 1 ... T keys and the
        // CTA waited for the longest), and the lanes of a warp read the same key row at the same time (broadcast).
        // Scores in the log2 domain: q is pre-multiplied by log2(e) / sqrt(hd) (attention.py:97-99 divides the scores by
        // sqrt(hd)), the softmax is ex2.approx(s - m) (2 ulp; the parity tests sit at 1e-6 against the 1e-5 bar).  One pass
        // over the keys in chunks of 8 with a running maximum (accumulators rescaled once per chunk).
        {
            const int pi = tid >> 2, h0 = (tid & 3) * HD;
#pragma unroll 1
            for (int side = 0; side < 2; side++) {
                const int t = side ? T - 1 - pi : pi;
                if (side ? (2 * pi >= T - 1) : (2 * pi > T - 1)) continue;    // tokens 0 .. T-1 once each (odd T: the middle one on side 0)
                const int nk = Tp + t + 1;
                float qv[HD];
#pragma unroll
                for (int d = 0; d < HD; d++) qv[d] = qs[t * FT_LDK + h0 + d] * (1.4426950408889634f / 2.8284271247461903f);
                float m0 = -INFINITY, den0 = 0.f, o[HD];
#pragma unroll
                for (int d = 0; d < HD; d++) o[d] = 0.f;
                for (int j0 = 0; j0 < nk; j0 += 8) {
                    float s0[8], c0 = -INFINITY;
#pragma unroll
                    for (int jj = 0; jj < 8; jj++) {
                        const float* kr = ks + min(j0 + jj, nk - 1) * FT_LDK + h0;
                        float u = 0.f;
#pragma unroll
                        for (int d = 0; d < HD; d++) u = fmaf(qv[d], kr[d], u);
                        s0[jj] = (j0 + jj < nk) ? u : -INFINITY;
                        c0 = fmaxf(c0, s0[jj]);
                    }
                    const float n0 = fmaxf(m0, c0);
                    const float r0 = ex2_approx(m0 - n0);      // 0 for the first chunk
                    m0 = n0;
                    den0 *= r0;
#pragma unroll
                    for (int d = 0; d < HD; d++) o[d] *= r0;
#pragma unroll
                    for (int jj = 0; jj < 8; jj++) {
                        const float e0 = ex2_approx(s0[jj] - m0);                  // masked keys: 2^-inf = 0
                        den0 += e0;
                        const float* vr = vs + min(j0 + jj, nk - 1) * FT_LDK + h0;
#pragma unroll
                        for (int d = 0; d < HD; d++) o[d] = fmaf(e0, vr[d], o[d]);
                    }
                }
                const float i0 = 1.0f / den0;
#pragma unroll
                for (int d = 0; d < HD; d++) a[t * FT_LDX + h0 + d] = o[d] * i0;
            }
        }
        __syncthreads();                        // all reads of this layer's keys / values are done
        ft_warp_gemm<E, E>(aw, FT_LDX, w + L.wproj, w + L.bproj, lane, [&](int r, int n, float v) { xw[r * FT_LDX + n] += v; });
        __syncwarp();
        ft_warp_ln(xw, aw, FT_LDX, w + L.ln2w, w + L.ln2b, p.eps, lane);
        __syncwarp();
        ft_warp_gemm<E, 4 * E>(aw, FT_LDX, w + L.wfc, w + L.bfc, lane, [&](int r, int n, float v) { hw[r * FT_LDH + n] = gelu_new_sig(v); });
        __syncwarp();
        ft_warp_gemm<4 * E, E>(hw, FT_LDH, w + L.wpr, w + L.bpr, lane, [&](int r, int n, float v) { xw[r * FT_LDX + n] += v; });
        __syncwarp();
    }
    ft_warp_ln(xw, aw, FT_LDX, p.blob + L.lnfw, p.blob + L.lnfb, p.eps, lane);
    __syncwarp();
    ft_warp_gemm<E, E>(aw, FT_LDX, p.blob + L.wf, p.blob + L.bf, lane, [&](int r, int n, float v) {
        const int t = r0 + r;
        if (t < T) p.hidden[((size_t)b * T + t) * E + n] = v;
    });
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    // out[c][r] = in[r][c]
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * cols) {
        const int r = i / cols, c = i - r * cols;
        out[(size_t)c * rows + r] = in[i];
    }
}


size_t generic_rollout_smem(const Gpt2Layout& L, int G) {
    return sizeof(float) * ((size_t)(L.E * G) * (1 + 1 + 3 + 4) + (size_t)4 * ROLL_THREADS * G +
                            (size_t)(ROLL_THREADS / 32) * L.n_ctx);
}

template <int G>
int launch_generic(trpx_gpt2_t h, const RolloutParams& p, int grid, size_t smem, cudaStream_t st) {
    TRPX_CUDA(cudaFuncSetAttribute(rollout_generic_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rollout_generic_kernel<G><<<grid, ROLL_THREADS, smem, st>>>(p);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

template <int G, int MAXI, bool CACHE>
int launch_e32_one(const RolloutParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    TRPX_CUDA(cudaFuncSetAttribute(rollout_e32_kernel<G, MAXI, CACHE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    rollout_e32_kernel<G, MAXI, CACHE><<<grid, block, smem, st>>>(p);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

template <int G>
int launch_e32(const RolloutParams& p, int maxi, int grid, int block, size_t smem, cudaStream_t st) {
    if (!p.use_cache) return launch_e32_one<G, 1, false>(p, grid, block, smem, st);
    switch (maxi) {
        case 1: return launch_e32_one<G, 1, true>(p, grid, block, smem, st);
        case 2: return launch_e32_one<G, 2, true>(p, grid, block, smem, st);
        case 4: return launch_e32_one<G, 4, true>(p, grid, block, smem, st);
        default: return launch_e32_one<G, 8, true>(p, grid, block, smem, st);
    }
}

template <int NCTX>
int launch_e32s_one(const RolloutParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    TRPX_CUDA(cudaFuncSetAttribute(rollout_e32s_kernel<NCTX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rollout_e32s_kernel<NCTX><<<grid, block, smem, st>>>(p);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

template <int NCTX, bool CACHE>
int launch_v2_one(const RolloutParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    TRPX_CUDA(cudaFuncSetAttribute(rollout_e32v2_kernel<NCTX, CACHE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    rollout_e32v2_kernel<NCTX, CACHE><<<grid, block, smem, st>>>(p);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

int launch_v2(const RolloutParams& p, int n_ctx, int grid, int block, size_t smem, cudaStream_t st) {
    if (!p.use_cache) return launch_v2_one<8, false>(p, grid, block, smem, st);
    if (n_ctx <= 8) return launch_v2_one<8, true>(p, grid, block, smem, st);
    if (n_ctx <= 16) return launch_v2_one<16, true>(p, grid, block, smem, st);
    if (n_ctx <= 32) return launch_v2_one<32, true>(p, grid, block, smem, st);
    return launch_v2_one<64, true>(p, grid, block, smem, st);
}

template <int NCTX, bool CACHE>
int launch_v4_one(const RolloutParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    TRPX_CUDA(cudaFuncSetAttribute(rollout_e32v4_kernel<NCTX, CACHE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    rollout_e32v4_kernel<NCTX, CACHE><<<grid, block, smem, st>>>(p);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

int launch_v4(const RolloutParams& p, int n_ctx, int grid, int block, size_t smem, cudaStream_t st) {
    if (!p.use_cache) return launch_v4_one<8, false>(p, grid, block, smem, st);
    if (n_ctx <= 8) return launch_v4_one<8, true>(p, grid, block, smem, st);
    if (n_ctx <= 16) return launch_v4_one<16, true>(p, grid, block, smem, st);
    if (n_ctx <= 32) return launch_v4_one<32, true>(p, grid, block, smem, st);
    return launch_v4_one<64, true>(p, grid, block, smem, st);
}

template <int NCTX>
int launch_ws_one(const RolloutParams& p, int grid, size_t smem, cudaStream_t st) {
    TRPX_CUDA(cudaFuncSetAttribute(rollout_e32ws_kernel<NCTX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rollout_e32ws_kernel<NCTX><<<grid, WS_THREADS, smem, st>>>(p);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

int ensure_kv(trpx_gpt2_t h, size_t bytes) {
    if (bytes <= h->kv_bytes) return TRPX_OK;
    if (h->kv) {
        TRPX_CUDA(cudaDeviceSynchronize());
        TRPX_CUDA(cudaFree(h->kv));
        h->kv = nullptr;
        h->kv_bytes = 0;
    }
    TRPX_CUDA(cudaMalloc(&h->kv, bytes));
    h->kv_bytes = bytes;
    return TRPX_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
namespace {
// images derived from the packed blob: XOR-swizzled copy for the tensor-core kernel, per-(layer, rank) slices for the
// cluster kernel
int derive_weight_images(trpx_gpt2_t h, cudaStream_t st) {
    const Gpt2Layout& L = h->lay;
    const int E = L.E;
    if (E == 32) {
        swizzle_weights_kernel<<<64, 256, 0, st>>>(h->blob, h->blob_sw, L);
        TRPX_CUDA(cudaGetLastError());
    }
    if (E % (CLS * 4) == 0) {
        const ClusterLayout cl = ClusterLayout::make(E);
        const size_t n = (size_t)(L.n_layer + 1) * CLS * ((size_t)cl.att_size + cl.mlp_size);
        if (h->img_cl == nullptr) TRPX_CUDA(cudaMalloc(&h->img_cl, n * sizeof(float)));
        cluster_image_kernel<<<256, 256, 0, st>>>(h->blob, h->img_cl, L, cl);
        TRPX_CUDA(cudaGetLastError());
    }
    return TRPX_OK;
}
}  // namespace

extern "C" {

int trpx_gpt2_create(const trpx_gpt2_config* cfg, int device, trpx_gpt2_t* out) {
    TRPX_REQUIRE(cfg != nullptr && out != nullptr, "trpx_gpt2_create: null argument");
    TRPX_REQUIRE(cfg->n_embd > 0 && cfg->n_embd % 32 == 0 && cfg->n_embd <= 512,
                 "n_embd=%d unsupported (multiple of 32, <= 512)", cfg->n_embd);
    TRPX_REQUIRE(cfg->n_head > 0 && cfg->n_embd % cfg->n_head == 0, "n_embd %% n_head != 0");
    TRPX_REQUIRE(cfg->n_layer > 0 && cfg->n_layer <= 64, "n_layer=%d unsupported", cfg->n_layer);
    TRPX_REQUIRE(cfg->n_ctx > 0 && cfg->n_ctx <= 1024, "n_ctx=%d unsupported", cfg->n_ctx);
    int ndev = 0;
    TRPX_CUDA(cudaGetDeviceCount(&ndev));
    TRPX_REQUIRE(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
    DeviceGuard guard(device);
    trpx_gpt2* h = new (std::nothrow) trpx_gpt2();
    TRPX_REQUIRE(h != nullptr, "out of host memory");
    h->cfg = *cfg;
    h->device = device;
    h->lay = Gpt2Layout::make(cfg->n_embd, cfg->n_layer, cfg->n_ctx, cfg->n_head);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (cudaMalloc(&h->blob, sizeof(float) * h->lay.total) != cudaSuccess ||
        cudaMalloc(&h->blob_sw, sizeof(float) * h->lay.total) != cudaSuccess ||
        cudaMalloc(&h->pe, sizeof(float) * cfg->n_ctx * cfg->n_embd) != cudaSuccess) {
        cudaFree(h->blob);
        cudaFree(h->blob_sw);
        delete h;
        return set_error(TRPX_E_CUDA, "cudaMalloc of the weight blob failed");
    }
    *out = h;
    return TRPX_OK;
}

int trpx_gpt2_destroy(trpx_gpt2_t h) {
    if (!valid(h)) return set_error(TRPX_E_STATE, "trpx_gpt2_destroy: invalid handle");
    DeviceGuard guard(h->device);
    cudaFree(h->blob);
    cudaFree(h->blob_sw);
    cudaFree(h->img_cl);
    cudaFree(h->pe);
    cudaFree(h->kv);
    cudaFree(h->train_ws);
    h->magic = 0;
    delete h;
    return TRPX_OK;
}

/* Weights as ONE packed blob in the layout of gpt2_layout.cuh (what trpx_gpt2_train_fwd_bwd's gradient uses): the
 * fused optimizer keeps its master copy in this layout and hands it over after every step. */
int trpx_gpt2_load_blob(trpx_gpt2_t h, const float* blob_dev, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "load the weights once with trpx_gpt2_load_weights first");
    TRPX_REQUIRE(blob_dev != nullptr, "null blob");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    TRPX_CUDA(cudaMemcpyAsync(h->blob, blob_dev, sizeof(float) * h->lay.total, cudaMemcpyDeviceToDevice, st));
    return derive_weight_images(h, st);
}

int trpx_gpt2_num_weight_tensors(trpx_gpt2_t h) { return valid(h) ? 12 * h->cfg.n_layer + 4 : -1; }

int trpx_gpt2_load_weights(trpx_gpt2_t h, const float* const* w, int n, const float* pe_dev, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    TRPX_REQUIRE(w != nullptr && n == 12 * h->cfg.n_layer + 4, "expected %d weight tensors, got %d",
                 12 * h->cfg.n_layer + 4, n);
    for (int i = 0; i < n; i++) TRPX_REQUIRE(w[i] != nullptr, "weight tensor %d is null", i);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const Gpt2Layout& L = h->lay;
    const int E = L.E;
    auto cp = [&](int off, const float* src, size_t cnt) {
        return cudaMemcpyAsync(h->blob + off, src, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st);
    };
    for (int l = 0; l < L.n_layer; l++) {
        const int base = l * L.layer_stride;
        const float* const* t = w + 12 * l;
        const int offs[12] = {L.ln1w, L.ln1b, L.wqkv, L.bqkv, L.wproj, L.bproj, L.ln2w, L.ln2b, L.wfc, L.bfc, L.wpr, L.bpr};
        const size_t cnts[12] = {(size_t)E, (size_t)E, (size_t)3 * E * E, (size_t)3 * E, (size_t)E * E, (size_t)E,
                                 (size_t)E, (size_t)E, (size_t)4 * E * E, (size_t)4 * E, (size_t)4 * E * E, (size_t)E};
        for (int i = 0; i < 12; i++) TRPX_CUDA(cp(base + offs[i], t[i], cnts[i]));
    }
    const float* const* g = w + 12 * L.n_layer;
    TRPX_CUDA(cp(L.lnfw, g[0], E));
    TRPX_CUDA(cp(L.lnfb, g[1], E));
    transpose_kernel<<<(E * E + 255) / 256, 256, 0, st>>>(g[2], h->blob + L.wf, E, E);   // [out][in] -> [in][out]
    TRPX_CUDA(cudaGetLastError());
    TRPX_CUDA(cp(L.bf, g[3], E));
    {
        int rc = derive_weight_images(h, st);
        if (rc) return rc;
    }
    if (pe_dev != nullptr) {
        TRPX_CUDA(cudaMemcpyAsync(h->pe, pe_dev, sizeof(float) * L.n_ctx * E, cudaMemcpyDeviceToDevice, st));
    } else {
        // phys_transformer_gpt2.py:215-219 evaluated in double, rounded once to float
        std::vector<float> pe((size_t)L.n_ctx * E);
        for (int p = 0; p < L.n_ctx; p++)
            for (int c = 0; c < E; c++) {
                if (c % 2 == 0) {
                    const float denom = powf(10000.0f, (float)(2.0f * (float)(c / 2) / (float)E));
                    pe[(size_t)p * E + c] = (float)sin((double)((float)p / denom));
                } else {
                    pe[(size_t)p * E + c] = (float)cos((double)p);
                }
            }
        TRPX_CUDA(cudaMemcpyAsync(h->pe, pe.data(), sizeof(float) * pe.size(), cudaMemcpyHostToDevice, st));
        TRPX_CUDA(cudaStreamSynchronize(st));
    }
    h->loaded = true;
    return TRPX_OK;
}

int trpx_gpt2_set_rollout_tuning(trpx_gpt2_t h, int variant, int warps_per_cta, int traj_per_warp, int max_ctas) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    TRPX_REQUIRE(variant >= 0 && variant <= 11, "variant must be 0..11");
    TRPX_REQUIRE(warps_per_cta >= 0 && warps_per_cta <= 32, "warps_per_cta must be 0..32");
    TRPX_REQUIRE(traj_per_warp == 0 || traj_per_warp == 1 || traj_per_warp == 2 || traj_per_warp == 4 ||
                     traj_per_warp == 8 || (traj_per_warp == 16 && (variant == 7 || variant == 9)),
                 "traj_per_warp must be 0,1,2,4,8 (16 with variant 7)");
    h->tune_variant = variant;
    h->tune_prefetch = (variant == 4) ? 0 : 1;
    h->tune_warps = warps_per_cta;
    h->tune_g = traj_per_warp;
    h->tune_max_ctas = max_ctas;
    return TRPX_OK;
}

int trpx_gpt2_last_launch(trpx_gpt2_t h, int* variant, int* grid, int* block, int* smem, int* group) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (variant) *variant = h->last_variant;
    if (grid) *grid = h->last_grid;
    if (block) *block = h->last_block;
    if (smem) *smem = h->last_smem;
    if (group) *group = h->last_group;
    return TRPX_OK;
}

int trpx_gpt2_rollout(trpx_gpt2_t h, const float* x0, long long x0_stride, float* out, long long out_stride, int B,
                      int S, int use_cache, float* presents, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "trpx_gpt2_rollout: weights not loaded");
    TRPX_REQUIRE(B >= 0 && S >= 0, "negative size");
    TRPX_REQUIRE(presents == nullptr || use_cache, "presents requires use_cache");
    if (B == 0 || S == 0) return TRPX_OK;
    TRPX_REQUIRE(x0 != nullptr && out != nullptr, "null tensor");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const Gpt2Layout& L = h->lay;
    RolloutParams p{};
    p.blob = h->blob; p.blob_sw = h->blob_sw; p.pe = h->pe; p.x0 = x0; p.x0_stride = x0_stride; p.out = out; p.out_stride = out_stride;
    p.presents = presents; p.B = B; p.S = S; p.use_cache = use_cache ? 1 : 0; p.eps = h->cfg.ln_eps; p.lay = L;
    p.use_tma = (h->tune_variant == 3) ? 0 : 1;
    p.prefetch = h->tune_prefetch;
    const size_t kv_per_traj = (size_t)L.n_layer * 2 * L.n_ctx * L.E * sizeof(float);

    // ---- E = 32 / head size 8 warp kernel, if the whole model fits in shared memory ----
    const bool e32_shape = (L.E == 32 && L.n_head == 4 && L.n_ctx <= 64);
    const size_t wbytes = (size_t)L.total * sizeof(float);
    const int scr_per_warp_g1 = 64 * (int)sizeof(float);
    // ---- tensor-core kernel (3xTF32 mma.sync, 16 trajectories per warp) ----
    {
        const int sms = h->tune_max_ctas > 0 ? std::min(h->tune_max_ctas, h->sm_count) : h->sm_count;
        const int ngroups = (B + V4G - 1) / V4G;
        const bool aligned = ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(out)) & 7) == 0 &&
                             (x0_stride % 2) == 0 && (out_stride % 2) == 0;
        const size_t scr_w = use_cache ? (size_t)V4G * V4QS * sizeof(float) : 0;
        const bool fits = wbytes + scr_w + 1024 <= (size_t)h->smem_optin;
        // measured (profiles/): without the KV window the tensor-core kernel is 1.7x the SIMT kernels from ~3.5 groups of 16
        // trajectories per SM on (B = 8,192: 411 against 333 M/s, 16,384: 540 against 312; B = 4,096: 208 against 257);
        // with the window the SIMT kernel hides the window latency better (more warps per trajectory)
        const bool want = h->tune_variant == 6 ||
                          (h->tune_variant == 0 && !use_cache && 4LL * ngroups >= 13LL * sms);
        if (e32_shape && aligned && fits && want) {
            const int max_w = scr_w ? (int)std::min<size_t>(8, ((size_t)h->smem_optin - 1024 - wbytes) / scr_w) : 12;
            const int grid = std::min(sms, ngroups);
            const int wantw = h->tune_warps > 0 ? h->tune_warps : 8;
            const int Wc = std::max(1, std::min(std::min(wantw, max_w), (ngroups + grid - 1) / grid));
            const size_t smem = wbytes + (size_t)Wc * scr_w;
            if (use_cache) {
                int rc = ensure_kv(h, (size_t)grid * Wc * V4G * kv_per_traj);
                if (rc) return rc;
            }
            p.kv = h->kv;
            p.prefetch = h->tune_prefetch ? 3 : 0;       // bit 0: first pass at layer start, bit 1: next pass during a pass
            if (const char* e = getenv("TRPX_V4_PF")) p.prefetch = atoi(e);      // tuning knob (scripts/sweep_rollout.py)
            h->last_variant = 6; h->last_grid = grid; h->last_block = Wc * 32; h->last_smem = (int)smem; h->last_group = V4G;
            return launch_v4(p, L.n_ctx, grid, Wc * 32, smem, st);
        }
    }
    // ---- warp-specialised streaming kernel (variant 11): dense warps + attention warps ----
    if (e32_shape && use_cache && h->tune_variant == 11 && (L.n_ctx == 8 || L.n_ctx == 16 || L.n_ctx == 32 || L.n_ctx == 64)) {
        const int sms = h->tune_max_ctas > 0 ? std::min(h->tune_max_ctas, h->sm_count) : h->sm_count;
        const int ngroups = (B + V4G - 1) / V4G;
        const bool aligned = ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(out)) & 7) == 0 &&
                             (x0_stride % 2) == 0 && (out_stride % 2) == 0;
        const size_t smem = wbytes + (size_t)WS_NG * V4G * V4QS * sizeof(float);
        if (aligned && smem + 1024 <= (size_t)h->smem_optin) {
            const int grid = std::min(sms, ngroups);
            int rc = ensure_kv(h, (size_t)grid * WS_NG * V4G * kv_per_traj);
            if (rc) return rc;
            p.kv = h->kv;
            h->last_variant = 11; h->last_grid = grid; h->last_block = WS_THREADS; h->last_smem = (int)smem; h->last_group = V4G;
            p.prefetch = 0x11;                            // low nibble: passes requested by the dense warp, high: attention look-ahead
            if (const char* e = getenv("TRPX_WS_PF")) p.prefetch = (int)strtol(e, nullptr, 0);      // tuning knob
            p.spin_ns = 100;
            if (const char* e = getenv("TRPX_WS_NS")) p.spin_ns = atoi(e);
#if WS_PROF
            static long long* prof_dev = nullptr;
            if (getenv("TRPX_WS_PROF")) {
                if (!prof_dev) TRPX_CUDA(cudaMalloc(&prof_dev, 64 * sizeof(long long)));
                p.prof = prof_dev;
                int rc2;
                switch (L.n_ctx) {
                    case 8: rc2 = launch_ws_one<8>(p, grid, smem, st); break;
                    case 16: rc2 = launch_ws_one<16>(p, grid, smem, st); break;
                    case 32: rc2 = launch_ws_one<32>(p, grid, smem, st); break;
                    default: rc2 = launch_ws_one<64>(p, grid, smem, st); break;
                }
                if (rc2) return rc2;
                long long hp[64];
                TRPX_CUDA(cudaStreamSynchronize(st));
                TRPX_CUDA(cudaMemcpy(hp, prof_dev, sizeof(hp), cudaMemcpyDeviceToHost));
                for (int w = 0; w < WS_DW + WS_NG; w++)
                    fprintf(stderr, "ws prof B=%d S=%d warp %2d (%s): wait %lld of %lld cycles (%.1f %%)\n", B, S, w, w < WS_DW ? "dense" : "attn ",
                            hp[2 * w], hp[2 * w + 1], 100.0 * (double)hp[2 * w] / (double)hp[2 * w + 1]);
                fprintf(stderr, "ws prof B=%d dense warp 0: proj+MLP %lld, LN1+QKV+stores %lld, fence+arrive %lld cycles\n", B, hp[32], hp[33], hp[34]);
                return TRPX_OK;
            }
#endif
            switch (L.n_ctx) {
                case 8: return launch_ws_one<8>(p, grid, smem, st);
                case 16: return launch_ws_one<16>(p, grid, smem, st);
                case 32: return launch_ws_one<32>(p, grid, smem, st);
                default: return launch_ws_one<64>(p, grid, smem, st);
            }
        }
    }
    // ---- second-generation kernel (16 trajectories per warp, 2-D lane tiling): large batches ----
    {
        const int sms = h->tune_max_ctas > 0 ? std::min(h->tune_max_ctas, h->sm_count) : h->sm_count;
        const int ngroups = (B + V2G - 1) / V2G;
        const bool aligned = ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                             (x0_stride % 4) == 0 && (out_stride % 4) == 0;
        const size_t scr_w = (size_t)32 * V2G * sizeof(float);
        const bool fits = wbytes + scr_w + 1024 <= (size_t)h->smem_optin;
        // measured (profiles/r01_sweep_rollout_v2.jsonl): v2 wins without the KV window, v1 (G = 4) with it
        const bool want = h->tune_variant == 5;
        if (e32_shape && aligned && fits && want) {
            const int max_w = (int)std::min<size_t>(12, ((size_t)h->smem_optin - 1024 - wbytes) / scr_w);
            const int grid = std::min(sms, ngroups);
            const int wantw = h->tune_warps > 0 ? h->tune_warps : max_w;
            const int Wc = std::max(1, std::min(std::min(wantw, max_w), (ngroups + grid - 1) / grid));
            const size_t smem = wbytes + (size_t)Wc * scr_w;
            if (use_cache) {
                int rc = ensure_kv(h, (size_t)grid * Wc * V2G * kv_per_traj);
                if (rc) return rc;
            }
            p.kv = h->kv;
            h->last_variant = 5; h->last_grid = grid; h->last_block = Wc * 32; h->last_smem = (int)smem; h->last_group = V2G;
            return launch_v2(p, L.n_ctx, grid, Wc * 32, smem, st);
        }
    }
    // ---- E=32, KV window on chip in tensor memory (variant 10, gpt2_tmem.cu) ----
    // Selected automatically in the latency-bound regime: while the batch fills only a few rounds of resident tiles
    // (4 rounds of 8 trajectories per SM for n_ctx <= 32, 2 rounds of 4 for n_ctx = 64) its 8 warps per tile finish a step
    // sooner than one warp per 1-8 trajectories does (measured: Lorenz 256 x 128 — BASELINE config 1 — 1.16 ms against
    // 2.18 ms; Rossler 1,184 x 128 1.10 against 2.35 ms, 4,736 x 128 4.29 against 4.52 ms; Lorenz 1,776 x 128 is the tie;
    // at full batch the streaming kernel is ahead, profiles/README.md).
    const bool tmem_ok = e32_shape && tmem_rollout_group(L) > 0 &&
                         tmem_rollout_smem(L) + 1024 <= (size_t)h->smem_optin &&
                         (reinterpret_cast<uintptr_t>(out) & 7) == 0 && (out_stride % 2) == 0 &&
                         (reinterpret_cast<uintptr_t>(presents) & 15) == 0;
    // use_cache=False (the reference's default mode) runs on the same cooperative tile without the window phases: ahead of
    // the warp-per-group kernels only while the batch is ONE round of tiles (Lorenz 256 x 128: 0.86 against 1.42 ms,
    // Rossler 1,184 x 128: 0.88 against 1.42 ms; two rounds already lose).
    const bool tmem_auto = tmem_ok && h->tune_variant == 0 && h->tune_g == 0 && h->tune_warps == 0 &&
                           (B + tmem_rollout_group(L) - 1) / tmem_rollout_group(L) <=
                               (!use_cache ? 1 : (L.n_ctx <= 32 ? 4 : 2)) *
                                   (h->tune_max_ctas > 0 ? std::min(h->tune_max_ctas, h->sm_count) : h->sm_count);
    if (tmem_ok && (h->tune_variant == 10 || tmem_auto)) {
        const int sms = h->tune_max_ctas > 0 ? std::min(h->tune_max_ctas, h->sm_count) : h->sm_count;
        const int G = tmem_rollout_group(L);
        const int ngroups = (B + G - 1) / G;
        const int grid = std::min(sms, ngroups);
        TmemRolloutParams tp{};
        tp.blob_sw = h->blob_sw; tp.pe = h->pe; tp.x0 = x0; tp.x0_stride = x0_stride; tp.out = out; tp.out_stride = out_stride;
        tp.presents = presents; tp.B = B; tp.S = S; tp.use_cache = use_cache ? 1 : 0; tp.eps = h->cfg.ln_eps; tp.lay = L;
        h->last_variant = 10; h->last_grid = grid; h->last_block = 256; h->last_smem = (int)tmem_rollout_smem(L); h->last_group = G;
        return launch_rollout_tmem(tp, grid, st);
    }
    // ---- E=32, KV window, TMA-staged attention + weight ring (variant 8) ----
    if (e32_shape && use_cache && (h->tune_variant == 8) && (L.n_ctx == 8 || L.n_ctx == 16 || L.n_ctx % 32 == 0)) {
        const int sms = h->tune_max_ctas > 0 ? std::min(h->tune_max_ctas, h->sm_count) : h->sm_count;
        const int nctx_t = L.n_ctx <= 8 ? 8 : (L.n_ctx <= 16 ? 16 : (L.n_ctx <= 32 ? 32 : 64));
        const int br = std::min(nctx_t, 32);
        const size_t per_warp = (size_t)(64 * E32S_G + E32S_G * br * 32) * sizeof(float);
        const size_t fixed = ((size_t)((L.total - L.lnfw + 3) & ~3) + 2 * (size_t)L.layer_stride) * sizeof(float);
        const int ngroups = (B + E32S_G - 1) / E32S_G;
        const int max_w = (int)std::min<size_t>(nctx_t >= 32 ? 7 : E32S_MAXW, ((size_t)h->smem_optin - 1024 - fixed) / per_warp);
        if (max_w >= 1) {
            const int grid = std::min(sms, ngroups);
            const int want = h->tune_warps > 0 ? h->tune_warps : max_w;
            const int Wc = std::max(1, std::min(std::min(want, max_w), (ngroups + grid - 1) / grid));
            const size_t smem = fixed + (size_t)Wc * per_warp;
            int rc = ensure_kv(h, (size_t)grid * Wc * E32S_G * kv_per_traj);
            if (rc) return rc;
            p.kv = h->kv;
            h->last_variant = 8; h->last_grid = grid; h->last_block = (Wc + 1) * 32; h->last_smem = (int)smem; h->last_group = E32S_G;
            switch (nctx_t) {
                case 8: return launch_e32s_one<8>(p, grid, (Wc + 1) * 32, smem, st);
                case 16: return launch_e32s_one<16>(p, grid, (Wc + 1) * 32, smem, st);
                case 32: return launch_e32s_one<32>(p, grid, (Wc + 1) * 32, smem, st);
                default: return launch_e32s_one<64>(p, grid, (Wc + 1) * 32, smem, st);
            }
        }
    }
    bool use_e32 = e32_shape && h->tune_variant != 1 && h->tune_variant != 7 && h->tune_variant != 9 && wbytes + (size_t)scr_per_warp_g1 + 1024 <= (size_t)h->smem_optin;
    if (use_e32) {
        const int sms = h->tune_max_ctas > 0 ? std::min(h->tune_max_ctas, h->sm_count) : h->sm_count;
        int G = h->tune_g;
        if (G == 0) {
            // measured (profiles/r01_sweep_*): with the KV window 4 trajectories per warp (two lanes per
            // (trajectory, head) pair, smaller in-flight window set) beat 8; without it 8 amortise the weights best.
            // Shrink the group further when the batch is too small to give every SM ~4 warps.
            G = (use_cache && L.n_ctx > 32) ? 4 : 8;     // n_ctx <= 32: 8 with 16 rows in flight measured 3 % ahead of 4
            while (G > 1 && (long long)(B + G - 1) / G < (long long)sms * 4) G >>= 1;
        }
        const int ngroups = (B + G - 1) / G;
        const int max_w = (int)std::min<size_t>((use_cache && G >= 4) ? TRPX_E32_CACHE_NT / 32 : 12,
                                                ((size_t)h->smem_optin - 1024 - wbytes) / ((size_t)64 * G * sizeof(float)));
        if (max_w < 1) {
            use_e32 = false;
        } else {
            const int grid = std::min(sms, ngroups);
            const int want = h->tune_warps > 0 ? h->tune_warps : ((use_cache && L.n_ctx <= 32) ? 11 : 8);
            int Wc = std::max(1, std::min(std::min(want, max_w), (ngroups + grid - 1) / grid));
            if (use_cache && h->tune_warps == 0) {
                // Wave quantisation: a warp runs its group at about the same speed whether the round is full or not, so the
                // pass takes `rounds` rounds; the FEWEST warps per CTA that still need only that many rounds make every
                // round faster (measured, Rossler G = 8: B = 12,288: 6 warps 22.8 ms against 25.1 with 8; 16,384: 7 warps
                // 24.3 against 25.9; 32,768: 49.1 against 52.6; 65,536 keeps 8)
                const long long slots = (long long)grid;
                const long long rounds = (ngroups + slots * Wc - 1) / (slots * Wc);
                while (Wc > 1 && (ngroups + slots * (Wc - 1) - 1) / (slots * (Wc - 1)) == rounds) Wc--;
            }
            const size_t smem = wbytes + (size_t)Wc * 64 * G * sizeof(float);
            if (use_cache) {
                int rc = ensure_kv(h, (size_t)grid * Wc * G * kv_per_traj);
                if (rc) return rc;
            }
            p.kv = h->kv;
            int maxi = 1;
            while (maxi * 8 < L.n_ctx) maxi <<= 1;
            h->last_variant = 2; h->last_grid = grid; h->last_block = Wc * 32; h->last_smem = (int)smem; h->last_group = G;
            switch (G) {
                case 8: return launch_e32<8>(p, maxi, grid, Wc * 32, smem, st);
                case 4: return launch_e32<4>(p, maxi, grid, Wc * 32, smem, st);
                case 2: return launch_e32<2>(p, maxi, grid, Wc * 32, smem, st);
                default: return launch_e32<1>(p, maxi, grid, Wc * 32, smem, st);
            }
        }
    }

    // ---- cluster kernel: wide model, few trajectories ----
    if (h->img_cl != nullptr && h->tune_variant != 1 && (!e32_shape || h->tune_variant == 7 || h->tune_variant == 9)) {
        const ClusterLayout cl = ClusterLayout::make(L.E);
        const int NTc = (h->tune_warps == 16) ? 512 : 256;
        auto smem_of = [&](int G) {
            return sizeof(float) * ((size_t)cl.att_size + cl.mlp_size + (size_t)L.E * G * ((h->tune_variant != 9 && G <= 2) ? 13 : 11) +
                                    (size_t)(NTc / 32) * 4 * cl.sl * G + (size_t)(NTc / 32) * L.n_ctx);
        };
        const bool async_x = h->tune_variant != 9;            // variant 9: cluster-barrier exchange (A/B reference)
        auto kernel_of = [&](int G) -> void (*)(ClusterParams) {
            if (async_x && G <= 2)
                return NTc == 512 ? (G == 2 ? rollout_cluster_async_kernel<2, 512> : rollout_cluster_async_kernel<1, 512>)
                                  : (G == 2 ? rollout_cluster_async_kernel<2, 256> : rollout_cluster_async_kernel<1, 256>);
#define TRPX_CLK(NT_)                                                                                         \
    (G == 16 ? rollout_cluster_kernel<16, NT_>                                                                \
             : G == 8 ? rollout_cluster_kernel<8, NT_>                                                        \
                      : G == 4 ? rollout_cluster_kernel<4, NT_>                                               \
                               : G == 2 ? rollout_cluster_kernel<2, NT_> : rollout_cluster_kernel<1, NT_>)
            return NTc == 512 ? TRPX_CLK(512) : TRPX_CLK(256);
#undef TRPX_CLK
        };
        int ql = 1;
        while (ql < cl.sl) ql <<= 1;                 // the FC slice has 4*sl columns = sl quads
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CLS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        // smallest group size whose clusters are all co-resident (one wave): a step costs about the same for
        // any group size (it is a chain of exchanges), so waves are what to avoid
        int G = 0, max_clusters = 0;
        for (int g = 1; g <= 16; g <<= 1) {
            if (h->tune_g && g != h->tune_g) continue;
            if (smem_of(g) + 1024 > (size_t)h->smem_optin || ql * g > NTc) break;
            cudaLaunchConfig_t q{};
            q.gridDim = dim3(CLS); q.blockDim = dim3(NTc); q.dynamicSmemBytes = smem_of(g);
            q.attrs = attr; q.numAttrs = 1;
            TRPX_CUDA(cudaFuncSetAttribute(kernel_of(g), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(g)));
            int nc = 0;
            TRPX_CUDA(cudaOccupancyMaxActiveClusters(&nc, kernel_of(g), &q));
            G = g; max_clusters = nc;
            // automatic selection also accepts TWO waves of single-trajectory clusters (2 x 51 us per step still beats
            // the generic kernel's 123 us)
            if ((B + g - 1) / g <= nc || (h->tune_variant == 0 && g == 1 && B <= 2 * nc)) break;
        }
        const int ngroups = G ? (B + G - 1) / G : 0;
        // measured (profiles/README.md): one wave of clusters of 1-2 trajectories (st.async exchange) runs a step in
        // 51-58 us against the generic kernel's 123 us; larger groups (cluster-barrier exchange) lose to it
        const bool want = G && max_clusters > 0 &&
                          (h->tune_variant == 7 || h->tune_variant == 9 ||
                           (h->tune_variant == 0 && G <= 2 && ngroups <= (G == 1 ? 2 : 1) * max_clusters));
        if (want) {
            int rc = ensure_kv(h, (size_t)ngroups * G * kv_per_traj);
            if (rc) return rc;
            ClusterParams cp{};
            cp.img = h->img_cl; cp.pe = h->pe; cp.x0 = x0; cp.x0_stride = x0_stride; cp.out = out; cp.out_stride = out_stride;
            cp.kv = h->kv; cp.presents = presents; cp.B = B; cp.S = S; cp.use_cache = use_cache ? 1 : 0;
            cp.eps = h->cfg.ln_eps; cp.lay = L; cp.cl = cl;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(ngroups * CLS);
            cfg.blockDim = dim3(NTc);
            cfg.dynamicSmemBytes = smem_of(G);
            cfg.stream = st;
            cfg.attrs = attr; cfg.numAttrs = 1;
            h->last_variant = h->tune_variant == 9 ? 9 : 7; h->last_grid = ngroups * CLS; h->last_block = NTc;
            h->last_smem = (int)smem_of(G); h->last_group = G;
            TRPX_CUDA(cudaLaunchKernelEx(&cfg, kernel_of(G), cp));
            return TRPX_OK;
        }
    }
    // ---- generic kernel ----
    int G = h->tune_g;
    if (G == 0) {
        // every CTA re-streams all weights each step, so few large groups amortise them; but a step is a latency chain,
        // and one trajectory per CTA is the fastest step (measured, Cylinder: 126 us against 153 us for G = 2 / 4 and
        // 282 us for G = 8) as long as every CTA gets its own SM
        G = 8;
        while (G > 1 && (B + G - 1) / G < h->sm_count / 2) G >>= 1;
        if (B <= h->sm_count) G = 1;
    }
    size_t smem = generic_rollout_smem(L, G);
    while (G > 1 && smem > (size_t)h->smem_optin) {
        G >>= 1;
        smem = generic_rollout_smem(L, G);
    }
    TRPX_REQUIRE(smem <= (size_t)h->smem_optin, "configuration needs %zu B of shared memory", smem);
    const int grid = (B + G - 1) / G;
    {
        int rc = ensure_kv(h, (size_t)grid * G * kv_per_traj);
        if (rc) return rc;
    }
    p.kv = h->kv;
    h->last_variant = 1; h->last_grid = grid; h->last_block = ROLL_THREADS; h->last_smem = (int)smem; h->last_group = G;
    switch (G) {
        case 8: return launch_generic<8>(h, p, grid, smem, st);
        case 4: return launch_generic<4>(h, p, grid, smem, st);
        case 2: return launch_generic<2>(h, p, grid, smem, st);
        default: return launch_generic<1>(h, p, grid, smem, st);
    }
}

int trpx_gpt2_forward(trpx_gpt2_t h, const float* x, const float* const* past, int Tp, float* hidden,
                      float* const* presents, float* const* attn, int B, int T, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "trpx_gpt2_forward: weights not loaded");
    TRPX_REQUIRE(B >= 0 && T >= 0 && Tp >= 0, "negative size");
    TRPX_REQUIRE((Tp == 0) == (past == nullptr), "past pointer / Tp mismatch");
    TRPX_REQUIRE(Tp + T <= h->cfg.n_ctx, "past (%d) + tokens (%d) exceeds n_ctx (%d)", Tp, T, h->cfg.n_ctx);
    if (B == 0 || T == 0) return TRPX_OK;
    TRPX_REQUIRE(x != nullptr && hidden != nullptr, "null tensor");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const Gpt2Layout& L = h->lay;
    const int nl = L.n_layer;
    TRPX_REQUIRE(nl <= MAX_FWD_LAYERS, "forward supports at most %d layers", MAX_FWD_LAYERS);
    ForwardParams p{};
    for (int l = 0; l < nl; l++) {
        TRPX_REQUIRE(!past || past[l], "past[%d] is null", l);
        TRPX_REQUIRE(!presents || presents[l], "presents[%d] is null", l);
        TRPX_REQUIRE(!attn || attn[l], "attn[%d] is null", l);
        p.past[l] = past ? past[l] : nullptr;
        p.presents[l] = presents ? presents[l] : nullptr;
        p.attn[l] = attn ? attn[l] : nullptr;
    }
    p.has_past = past != nullptr; p.has_presents = presents != nullptr; p.has_attn = attn != nullptr;
    p.blob = h->blob; p.pe = h->pe; p.x = x; p.hidden = hidden;
    p.B = B; p.T = T; p.Tp = Tp; p.eps = h->cfg.ln_eps; p.lay = L;
    // E = 32 / 4 heads / window <= 64 without attention-weight output: tensor-core kernel (3xTF32 mma.sync)
    if (L.E == 32 && L.n_head == 4 && Tp + T <= FT_ROWS && attn == nullptr && h->tune_variant != 1) {
        const size_t smem_tc = sizeof(float) * ((size_t)FT_ROWS * (2 * FT_LDX + FT_LDH));
        TRPX_CUDA(cudaFuncSetAttribute(forward_e32_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
        forward_e32_tc_kernel<<<B, FT_THREADS, smem_tc, st>>>(p);
        TRPX_CUDA(cudaGetLastError());
        return TRPX_OK;
    }
    const int TPd = ((T + TG - 1) / TG) * TG;
    const size_t smem = sizeof(float) * ((size_t)TPd * L.E * 3 + (size_t)2 * L.n_ctx * L.E + (size_t)TPd * 4 * L.E +
                                         (size_t)GEN_THREADS * TG + (size_t)(GEN_THREADS / 32) * L.n_ctx);
    TRPX_REQUIRE(smem <= (size_t)h->smem_optin, "forward needs %zu B of shared memory", smem);
    TRPX_CUDA(cudaFuncSetAttribute(forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    forward_kernel<<<B, GEN_THREADS, smem, st>>>(p);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

}  // extern "C"
