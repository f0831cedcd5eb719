// Cylinder Koopman embedding (K5 conv encoder, K6 upsample+conv decoder, K4 per-sample banded Koopman).
// Reference: trphysx/embedding/embedding_cylinder.py:42-67 (encoder), :70-88 (decoder), :138-170
// (embed / recover), :172-196 (koopmanOperation).  All convs are 3x3 cross-correlations with
// padding=1, padding_mode='replicate'; upsampling is bilinear x2 with align_corners=True.
#include "cyl.cuh"
#include <cmath>
#include <vector>
#include <algorithm>

using namespace trpx;

namespace trpx {
void cyl_tensor_sizes(int E, size_t (&n)[36]) {
    const int c4 = E / 32;
    const int nband = (E - 1) + (E - 2) + (E - 3) + (E - 4);
    size_t s[36] = {4, 4,
                    16 * 4 * 9, 16, 32 * 16 * 9, 32, 64 * 32 * 9, 64, 128 * 64 * 9, 128, (size_t)c4 * 128 * 9, (size_t)c4,
                    (size_t)E, (size_t)E,
                    (size_t)128 * c4 * 9, 128, 64 * 128 * 9, 64, 32 * 64 * 9, 32, 16 * 32 * 9, 16, 3 * 16 * 9, 3,
                    50, 50, (size_t)E * 50, (size_t)E,
                    50, 50, (size_t)nband * 50, (size_t)nband,
                    50, 50, (size_t)nband * 50, (size_t)nband};
    for (int i = 0; i < 36; i++) n[i] = s[i];
}
}  // namespace trpx

namespace {

bool valid(trpx_cyl_t h) { return cyl_valid(h); }
void tensor_sizes(int E, size_t (&n)[36]) { cyl_tensor_sizes(E, n); }

// ---- encoder input: cat(x, visc) then (x - mu) / std  (embedding_cylinder.py:149-150, 198-200) ----------
__global__ void cyl_prep_kernel(const float* __restrict__ x, const float* __restrict__ visc, const float* __restrict__ mu,
                                const float* __restrict__ sd, float* __restrict__ out, int N) {
    const int HW = CYL_H * CYL_W;
    const long long total = (long long)N * 4 * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int pix = (int)(i % HW);
        const int c = (int)((i / HW) % 4);
        const int n = (int)(i / (4LL * HW));
        const float v = (c < 3) ? x[((long long)n * 3 + c) * HW + pix] : visc[n] * 1.0f;
        out[i] = (v - __ldg(mu + c)) / __ldg(sd + c);
    }
}

// ---- simple direct conv (small layers): one thread per output element ------------------------------------
// in [N][Cin][Hs][Ws]; if UP the conv runs on the x2 bilinear upsampling of `in`.  out [N][Cout][Ho][Wo].
template <bool UP>
__global__ void conv3x3_naive_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                     const float* __restrict__ bias, float* __restrict__ out, int N, int Cin, int Hs,
                                     int Ws, int Cout, int stride, int relu) {
    const int Hi = UP ? 2 * Hs : Hs, Wi = UP ? 2 * Ws : Ws;       // conv input size
    const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
    const long long total = (long long)N * Cout * Ho * Wo;
    const float sh = UP ? (float)(Hs - 1) / (float)(Hi - 1) : 0.f;
    const float sw = UP ? (float)(Ws - 1) / (float)(Wi - 1) : 0.f;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int xo = (int)(idx % Wo);
        const int yo = (int)((idx / Wo) % Ho);
        const int co = (int)((idx / ((long long)Wo * Ho)) % Cout);
        const int n = (int)(idx / ((long long)Wo * Ho * Cout));
        float acc = __ldg(bias + co);
        const float* wp = w + (size_t)co * Cin * 9;
        const float* ip = in + (size_t)n * Cin * Hs * Ws;
        for (int ci = 0; ci < Cin; ci++) {
            const float* ic = ip + (size_t)ci * Hs * Ws;
#pragma unroll
            for (int dy = 0; dy < 3; dy++) {
                const int yy = min(max(yo * stride + dy - 1, 0), Hi - 1);       // replicate padding
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    const int xx = min(max(xo * stride + dx - 1, 0), Wi - 1);
                    float v;
                    if (UP) {
                        const Lerp ly = lerp_idx(yy, Hs, sh), lx = lerp_idx(xx, Ws, sw);
                        const float v00 = ic[ly.i0 * Ws + lx.i0], v01 = ic[ly.i0 * Ws + lx.i1];
                        const float v10 = ic[ly.i1 * Ws + lx.i0], v11 = ic[ly.i1 * Ws + lx.i1];
                        v = ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
                    } else {
                        v = ic[yy * Ws + xx];
                    }
                    acc = fmaf(v, __ldg(wp + ci * 9 + dy * 3 + dx), acc);
                }
            }
        }
        if (relu) acc = fmaxf(acc, 0.f);
        out[idx] = acc;
    }
}

// ---- tiled direct conv for the heavy decoder layers (stride 1) -----------------------------------------
// CTA = (frame, 16x32 output tile, 2*CT output channels), 256 threads: thread = 4 x-adjacent pixels x CT
// channels.  Per chunk of CC input channels the CTA builds the (optionally upsampled, replicate-padded)
// input patch and the weight slice in shared memory; weights are read as warp-uniform broadcasts.
constexpr int TH = 16, TW = 32, PW = 36, CC = 8, CONV_THREADS = 256;   // static smem stays < 48 KB

struct ConvEpilogue {
    int relu;
    int unnorm;                     // y = sd[c]*y + mu[c]   (embedding_cylinder.py:202-203)
    const float* mu;
    const float* sd;
    const unsigned char* mask;      // nullptr or [H][W]: zero where set (:168-169)
};

template <int CT, bool UP>
__global__ void __launch_bounds__(CONV_THREADS) conv3x3_tiled_kernel(const float* __restrict__ in,
                                                                     const float* __restrict__ w,
                                                                     const float* __restrict__ bias,
                                                                     float* __restrict__ out, int Cin, int Hs, int Ws,
                                                                     int Cout, ConvEpilogue ep) {
    constexpr int COT = 2 * CT;
    __shared__ __align__(16) float patch[CC][TH + 2][PW];
    __shared__ __align__(16) float wsm[CC][9][COT];
    __shared__ Lerp ylut[TH + 2], xlut[TW + 2];
    const int H = UP ? 2 * Hs : Hs, W = UP ? 2 * Ws : Ws;
    const int tiles_x = (W + TW - 1) / TW;
    const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TW;
    const int co0 = blockIdx.y * COT;
    const int n = blockIdx.z;
    const int tid = threadIdx.x;
    const int grp = tid >> 7;                 // channel group (warp-uniform)
    const int pos = tid & 127;
    const int py = pos >> 3, px = (pos & 7) * 4;

    // per-tile lookup of the source rows/cols feeding each patch row/col
    if (tid < TH + 2) {
        const int yy = min(max(ty0 + tid - 1, 0), H - 1);
        Lerp l;
        if (UP) l = lerp_idx(yy, Hs, (float)(Hs - 1) / (float)(H - 1));
        else { l.i0 = yy; l.i1 = yy; l.l0 = 1.f; l.l1 = 0.f; }
        ylut[tid] = l;
    } else if (tid >= 64 && tid < 64 + TW + 2) {
        const int t = tid - 64;
        const int xx = min(max(tx0 + t - 1, 0), W - 1);
        Lerp l;
        if (UP) l = lerp_idx(xx, Ws, (float)(Ws - 1) / (float)(W - 1));
        else { l.i0 = xx; l.i1 = xx; l.l0 = 1.f; l.l1 = 0.f; }
        xlut[t] = l;
    }

    float acc[4][CT];
#pragma unroll
    for (int c = 0; c < CT; c++) {
        const int co = co0 + grp * CT + c;
        const float b = (co < Cout) ? __ldg(bias + co) : 0.f;
#pragma unroll
        for (int p = 0; p < 4; p++) acc[p][c] = b;
    }
    const float* ip = in + (size_t)n * Cin * Hs * Ws;
    __syncthreads();

    for (int c0 = 0; c0 < Cin; c0 += CC) {
        // weights slice: wsm[ci][tap][co]
        for (int i = tid; i < CC * 9 * COT; i += CONV_THREADS) {
            const int co = i % COT, r = i / COT;
            const int tap = r % 9, ci = r / 9;
            const int gco = co0 + co, gci = c0 + ci;
            wsm[ci][tap][co] = (gco < Cout && gci < Cin) ? __ldg(w + ((size_t)gco * Cin + gci) * 9 + tap) : 0.f;
        }
        // input patch
        for (int i = tid; i < CC * (TH + 2) * (TW + 2); i += CONV_THREADS) {
            const int xx = i % (TW + 2), r = i / (TW + 2);
            const int yy = r % (TH + 2), ci = r / (TH + 2);
            float v = 0.f;
            if (c0 + ci < Cin) {
                const float* ic = ip + (size_t)(c0 + ci) * Hs * Ws;
                const Lerp ly = ylut[yy], lx = xlut[xx];
                if (UP) {
                    const float v00 = __ldg(ic + ly.i0 * Ws + lx.i0), v01 = __ldg(ic + ly.i0 * Ws + lx.i1);
                    const float v10 = __ldg(ic + ly.i1 * Ws + lx.i0), v11 = __ldg(ic + ly.i1 * Ws + lx.i1);
                    v = ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
                } else {
                    v = __ldg(ic + ly.i0 * Ws + lx.i0);
                }
            }
            patch[ci][yy][xx] = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int ci = 0; ci < CC; ci++) {
#pragma unroll
            for (int dy = 0; dy < 3; dy++) {
                const float* row = &patch[ci][py + dy][px];
                const float4 a = *reinterpret_cast<const float4*>(row);
                const float2 b = *reinterpret_cast<const float2*>(row + 4);
                const float iv[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    float wv[CT];
                    const float* wp = &wsm[ci][dy * 3 + dx][grp * CT];
                    if constexpr (CT % 4 == 0) {
#pragma unroll
                        for (int j = 0; j < CT / 4; j++) {
                            const float4 t = *reinterpret_cast<const float4*>(wp + 4 * j);
                            wv[4 * j] = t.x; wv[4 * j + 1] = t.y; wv[4 * j + 2] = t.z; wv[4 * j + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CT; j++) wv[j] = wp[j];
                    }
#pragma unroll
                    for (int p = 0; p < 4; p++)
#pragma unroll
                        for (int c = 0; c < CT; c++) acc[p][c] = fmaf(iv[p + dx], wv[c], acc[p][c]);
                }
            }
        }
        __syncthreads();
    }

    const int y = ty0 + py, x = tx0 + px;
    if (y < H && x < W) {
#pragma unroll
        for (int c = 0; c < CT; c++) {
            const int co = co0 + grp * CT + c;
            if (co >= Cout) continue;
            float v[4];
#pragma unroll
            for (int p = 0; p < 4; p++) {
                float t = acc[p][c];
                if (ep.relu) t = fmaxf(t, 0.f);
                if (ep.unnorm) t = __ldg(ep.sd + co) * t + __ldg(ep.mu + co);
                if (ep.mask && x + p < W && ep.mask[y * W + x + p]) t = 0.f;
                v[p] = t;
            }
            float* op = out + (((size_t)n * Cout + co) * H + y) * W + x;
            if (x + 3 < W) *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
            else
                for (int p = 0; p < 4 && x + p < W; p++) op[p] = v[p];
        }
    }
}

// ---- tensor-core implicit GEMM for the heavy decoder layers (bilinear x2 upsample + 3x3 conv + ReLU) ------
// M = pixels, N = output channels, K = 9 taps x Cin, computed as 3xTF32 (a_lo*w_hi + a_hi*w_lo + a_hi*w_hi,
// FP32 accumulate) on mma.sync.m16n8k8 so the FP32 parity bar holds.  Per chunk of 8 input channels the CTA
// builds the upsampled, replicate-padded patch ONCE, already split into TF32 hi / lo planes, plus the hi / lo
// weight slice; a warp owns 2 output rows x 32 pixels (4 M-tiles) and NT*8 output channels.  Fragment loads are
// bank-conflict free: the patch plane stride (18*36 = 648 = 8 mod 32) and the padded weight row (9*(COUT+8) = 8
// or 24 mod 32) spread the four k-lanes q over disjoint bank groups.
constexpr int MCC = 8;                       // input channels per chunk = one k-step per tap
constexpr int MPLANE = (TH + 2) * PW;        // 648 floats

__device__ __forceinline__ void cyl_split_tf32(float x, float& hi, float& lo) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    hi = __uint_as_float(u);
    lo = x - hi;
}
__device__ __forceinline__ void cyl_mma(float (&c)[4], const float (&a)[4], float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

template <int NT>
__global__ void __launch_bounds__(CONV_THREADS) conv3x3_up_mma_kernel(const float* __restrict__ in,
                                                                      const float* __restrict__ w,
                                                                      const float* __restrict__ bias,
                                                                      float* __restrict__ out, int Cin, int Hs, int Ws,
                                                                      int Cout) {
    constexpr int COT = NT * 8;              // output channels per CTA
    constexpr int COP = COT + 8;             // padded weight row
    extern __shared__ __align__(16) float smc[];
    float* p_hi = smc;                       // [MCC][TH+2][PW]
    float* p_lo = p_hi + MCC * MPLANE;
    float* w_hi = p_lo + MCC * MPLANE;       // [MCC][9][COP]
    float* w_lo = w_hi + MCC * 9 * COP;
    __shared__ Lerp ylut[TH + 2], xlut[TW + 2];
    const int H = 2 * Hs, W = 2 * Ws;
    const int tiles_x = (W + TW - 1) / TW;
    const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TW;
    const int co0 = blockIdx.y * COT;
    const int n = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;

    if (tid < TH + 2) {
        const int yy = min(max(ty0 + tid - 1, 0), H - 1);
        ylut[tid] = lerp_idx(yy, Hs, (float)(Hs - 1) / (float)(H - 1));
    } else if (tid >= 64 && tid < 64 + TW + 2) {
        const int t = tid - 64;
        const int xx = min(max(tx0 + t - 1, 0), W - 1);
        xlut[t] = lerp_idx(xx, Ws, (float)(Ws - 1) / (float)(W - 1));
    }
    // accumulators: M-tile m = (row 2*warp + m/2, x0 = 16*(m%2)); C fragment rows g, g+8 = pixels, cols = channels
    float acc[4][NT][4];
#pragma unroll
    for (int t = 0; t < NT; t++) {
        const int co = co0 + 8 * t + 2 * q;
        const float b0 = co < Cout ? __ldg(bias + co) : 0.f, b1 = co + 1 < Cout ? __ldg(bias + co + 1) : 0.f;
#pragma unroll
        for (int m = 0; m < 4; m++) { acc[m][t][0] = b0; acc[m][t][1] = b1; acc[m][t][2] = b0; acc[m][t][3] = b1; }
    }
    const float* ip = in + (size_t)n * Cin * Hs * Ws;
    __syncthreads();

    for (int c0 = 0; c0 < Cin; c0 += MCC) {
        for (int i = tid; i < MCC * 9 * COT; i += CONV_THREADS) {
            const int co = i % COT, r = i / COT;
            const int tap = r % 9, ci = r / 9;
            const int gco = co0 + co, gci = c0 + ci;
            const float v = (gco < Cout && gci < Cin) ? __ldg(w + ((size_t)gco * Cin + gci) * 9 + tap) : 0.f;
            float hi, lo;
            cyl_split_tf32(v, hi, lo);
            w_hi[(ci * 9 + tap) * COP + co] = hi;
            w_lo[(ci * 9 + tap) * COP + co] = lo;
        }
        for (int i = tid; i < MCC * (TH + 2) * (TW + 2); i += CONV_THREADS) {
            const int xx = i % (TW + 2), r = i / (TW + 2);
            const int yy = r % (TH + 2), ci = r / (TH + 2);
            float v = 0.f;
            if (c0 + ci < Cin) {
                const float* ic = ip + (size_t)(c0 + ci) * Hs * Ws;
                const Lerp ly = ylut[yy], lx = xlut[xx];
                const float v00 = __ldg(ic + ly.i0 * Ws + lx.i0), v01 = __ldg(ic + ly.i0 * Ws + lx.i1);
                const float v10 = __ldg(ic + ly.i1 * Ws + lx.i0), v11 = __ldg(ic + ly.i1 * Ws + lx.i1);
                v = ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
            }
            float hi, lo;
            cyl_split_tf32(v, hi, lo);
            p_hi[ci * MPLANE + yy * PW + xx] = hi;
            p_lo[ci * MPLANE + yy * PW + xx] = lo;
        }
        __syncthreads();
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) {
                const int tap = dy * 3 + dx;
                float bh[NT][2], bl[NT][2];
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    const int o0 = (q * 9 + tap) * COP + 8 * t + g, o1 = ((q + 4) * 9 + tap) * COP + 8 * t + g;
                    bh[t][0] = w_hi[o0]; bh[t][1] = w_hi[o1];
                    bl[t][0] = w_lo[o0]; bl[t][1] = w_lo[o1];
                }
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int py = 2 * warp + (m >> 1) + dy, px = 16 * (m & 1) + g + dx;
                    const int a0 = q * MPLANE + py * PW + px, a2 = (q + 4) * MPLANE + py * PW + px;
                    const float ah[4] = {p_hi[a0], p_hi[a0 + 8], p_hi[a2], p_hi[a2 + 8]};
                    const float al[4] = {p_lo[a0], p_lo[a0 + 8], p_lo[a2], p_lo[a2 + 8]};
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        cyl_mma(acc[m][t], al, bh[t][0], bh[t][1]);
                        cyl_mma(acc[m][t], ah, bl[t][0], bl[t][1]);
                        cyl_mma(acc[m][t], ah, bh[t][0], bh[t][1]);
                    }
                }
            }
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int y = ty0 + 2 * warp + (m >> 1);
        if (y >= H) continue;
#pragma unroll
        for (int t = 0; t < NT; t++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int co = co0 + 8 * t + 2 * q + (i & 1);
                const int x = tx0 + 16 * (m & 1) + g + 8 * (i >> 1);
                if (co < Cout && x < W) out[(((size_t)n * Cout + co) * H + y) * W + x] = fmaxf(acc[m][t][i], 0.f);
            }
    }
}

// ---- pipelined tensor-core decoder layer (conv mode 2) -----------------------------------------------------
// Stage A (cyl_upsplit_kernel): bilinear x2 upsample + replicate pad + TF32 hi/lo split of a whole layer input,
//   written ONCE to an L2-resident scratch U[plane][n][ci][H+2][W+4]  (plane 0 = hi, 1 = lo).
// Stage B (conv3x3_pipe_mma_kernel): 3xTF32 implicit GEMM whose patch / weight tiles are plain sub-blocks of U
//   and of a pre-arranged weight image, fetched with 16-byte cp.async (LDGSTS) into a DOUBLE-BUFFERED shared
//   memory stage: no arithmetic and no exposed latency in the load phase.  Each 8-channel chunk is accumulated in
//   a zeroed temporary (27 MMAs) and added to the FP32 master accumulator with a rounded FADD, which removes the
//   truncation bias of long tensor-core accumulation chains (rel-L2 6e-6 -> ~1e-6).
template <bool UP>
__global__ void cyl_upsplit_kernel(const float* __restrict__ in, float* __restrict__ U, int NC, int Hs, int Ws,
                                   size_t plane_stride) {
    // one thread = 4 consecutive padded columns of one padded row: row interpolation set up once, 128-bit stores
    const int H = UP ? 2 * Hs : Hs, W = UP ? 2 * Ws : Ws, HP = H + 2, WP = W + 4, WQ = WP >> 2;
    const float sh = UP ? (float)(Hs - 1) / (float)(H - 1) : 0.f, sw = UP ? (float)(Ws - 1) / (float)(W - 1) : 0.f;
    const size_t total = (size_t)NC * HP * WQ;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int xq = (int)(i % WQ);
        const size_t row = i / WQ;                       // nc * HP + yp
        const int yp = (int)(row % HP);
        const size_t nc = row / HP;
        const int y = min(max(yp - 1, 0), H - 1);
        Lerp ly;
        if (UP) ly = lerp_idx(y, Hs, sh);
        else { ly.i0 = y; ly.i1 = y; ly.l0 = 1.f; ly.l1 = 0.f; }
        const float* r0 = in + nc * Hs * Ws + ly.i0 * Ws;
        const float* r1 = in + nc * Hs * Ws + ly.i1 * Ws;
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int x = min(max(4 * xq + j - 1, 0), W - 1);
            float v;
            if (UP) {
                const Lerp lx = lerp_idx(x, Ws, sw);
                v = ly.l0 * (lx.l0 * __ldg(r0 + lx.i0) + lx.l1 * __ldg(r0 + lx.i1)) +
                    ly.l1 * (lx.l0 * __ldg(r1 + lx.i0) + lx.l1 * __ldg(r1 + lx.i1));
            } else {
                v = __ldg(r0 + x);
            }
            cyl_split_tf32(v, hi[j], lo[j]);
        }
        float* dst = U + row * WP + 4 * xq;
        *reinterpret_cast<float4*>(dst) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(dst + plane_stride) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// weight image for the pipelined kernel: [pass][chunk][plane][ci 8][tap 9][COP], COP = NT*8 + 8 (zero padded)
__global__ void cyl_wimage_kernel(const float* __restrict__ w, float* __restrict__ img, int Cin, int Cout, int COT) {
    const int COP = COT + 8, nchunk = Cin / MCC, npass = (Cout + COT - 1) / COT;
    const int total = npass * nchunk * 2 * MCC * 9 * COP;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int r = i;
        const int co = r % COP; r /= COP;
        const int tap = r % 9; r /= 9;
        const int ci = r % MCC; r /= MCC;
        const int plane = r % 2; r /= 2;
        const int chunk = r % nchunk; const int pass = r / nchunk;
        const int gco = pass * COT + co, gci = chunk * MCC + ci;
        float v = 0.f;
        if (co < COT && gco < Cout) v = w[((size_t)gco * Cin + gci) * 9 + tap];
        float hi, lo;
        cyl_split_tf32(v, hi, lo);
        img[i] = plane ? lo : hi;
    }
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NT>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv3x3_pipe_mma_kernel(const float* __restrict__ U,
                                                                           size_t plane_stride,
                                                                           const float* __restrict__ wimg,
                                                                           const float* __restrict__ bias,
                                                                           float* __restrict__ out, int Cin, int H,
                                                                           int W, int Cout, ConvEpilogue ep) {
    constexpr int COT = NT * 8, COP = COT + 8;
    constexpr int PATCH = MCC * MPLANE;                 // floats per plane per stage
    constexpr int WSZ = MCC * 9 * COP;
    constexpr int STAGE = 2 * PATCH + 2 * WSZ;          // hi patch | lo patch | hi weights | lo weights
    extern __shared__ __align__(16) float smc[];
    const int HP = H + 2, WP = W + 4;
    const int tiles_x = W / TW;
    const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TW;
    const int pass = blockIdx.y, co0 = pass * COT;
    const int n = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int nchunk = Cin / MCC;

    auto issue = [&](int chunk, int buf) {
        float* st = smc + buf * STAGE;
        // patch rows: MCC channels x (TH+2) rows x 2 planes, 9 x 16 B per row
        for (int i = tid; i < 2 * MCC * (TH + 2) * 9; i += CONV_THREADS) {
            const int j = i % 9;
            int r = i / 9;
            const int row = r % (TH + 2); r /= (TH + 2);
            const int ci = r % MCC, plane = r / MCC;
            const float* src = U + plane * plane_stride +
                               (((size_t)n * Cin + chunk * MCC + ci) * HP + ty0 + row) * WP + tx0 + 4 * j;
            cp_async16(st + plane * PATCH + ci * MPLANE + row * PW + 4 * j, src);
        }
        const float* wsrc = wimg + ((size_t)pass * nchunk + chunk) * 2 * WSZ;
        for (int i = tid; i < 2 * WSZ / 4; i += CONV_THREADS) cp_async16(st + 2 * PATCH + 4 * i, wsrc + 4 * i);
        cp_async_commit();
    };

    float acc[4][NT][4];
#pragma unroll
    for (int t = 0; t < NT; t++) {
        const int co = co0 + 8 * t + 2 * q;
        const float b0 = co < Cout ? __ldg(bias + co) : 0.f, b1 = co + 1 < Cout ? __ldg(bias + co + 1) : 0.f;
#pragma unroll
        for (int m = 0; m < 4; m++) { acc[m][t][0] = b0; acc[m][t][1] = b1; acc[m][t][2] = b0; acc[m][t][3] = b1; }
    }
    issue(0, 0);
    for (int c = 0; c < nchunk; c++) {
        if (c + 1 < nchunk) {
            issue(c + 1, (c + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* st = smc + (c & 1) * STAGE;
        const float* p_hi = st;
        const float* p_lo = st + PATCH;
        const float* w_hi = st + 2 * PATCH;
        const float* w_lo = w_hi + WSZ;
        float tmp[4][NT][4];
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
            for (int t = 0; t < NT; t++) tmp[m][t][0] = tmp[m][t][1] = tmp[m][t][2] = tmp[m][t][3] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) {
                const int tap = dy * 3 + dx;
                float bh[NT][2], bl[NT][2];
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    const int o0 = (q * 9 + tap) * COP + 8 * t + g, o1 = ((q + 4) * 9 + tap) * COP + 8 * t + g;
                    bh[t][0] = w_hi[o0]; bh[t][1] = w_hi[o1];
                    bl[t][0] = w_lo[o0]; bl[t][1] = w_lo[o1];
                }
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int py = 2 * warp + (m >> 1) + dy, px = 16 * (m & 1) + g + dx;
                    const int a0 = q * MPLANE + py * PW + px, a2 = (q + 4) * MPLANE + py * PW + px;
                    const float ah[4] = {p_hi[a0], p_hi[a0 + 8], p_hi[a2], p_hi[a2 + 8]};
                    const float al[4] = {p_lo[a0], p_lo[a0 + 8], p_lo[a2], p_lo[a2 + 8]};
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        cyl_mma(tmp[m][t], al, bh[t][0], bh[t][1]);
                        cyl_mma(tmp[m][t], ah, bl[t][0], bl[t][1]);
                        cyl_mma(tmp[m][t], ah, bh[t][0], bh[t][1]);
                    }
                }
            }
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
            for (int t = 0; t < NT; t++)
#pragma unroll
                for (int i = 0; i < 4; i++) acc[m][t][i] += tmp[m][t][i];
        __syncthreads();                               // stage (c & 1) may be refilled by the next issue
    }
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int y = ty0 + 2 * warp + (m >> 1);
#pragma unroll
        for (int t = 0; t < NT; t++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int co = co0 + 8 * t + 2 * q + (i & 1);
                const int x = tx0 + 16 * (m & 1) + g + 8 * (i >> 1);
                if (co < Cout) {
                    float v = acc[m][t][i];
                    if (ep.relu) v = fmaxf(v, 0.f);
                    if (ep.unnorm) v = __ldg(ep.sd + co) * v + __ldg(ep.mu + co);
                    if (ep.mask && ep.mask[y * W + x]) v = 0.f;
                    out[(((size_t)n * Cout + co) * H + y) * W + x] = v;
                }
            }
    }
}

// ---- first decoder layer: g.view(4,4,8) -> bilinear x2 -> conv(4 -> 128) -> ReLU, one CTA per frame ---------
// (1.2 MFLOP per frame: the generic direct kernel spent more time here than in a 75 MFLOP layer.)  Thread =
// (pixel, half of the output channels): its 4x3x3 input window sits in 36 registers, weights are warp-uniform
// LDS.128 broadcasts, stores are coalesced over pixels.
template <bool CL4>       // CL4: output in channels-last-4 order [n][CO/4][pixel][4] for the UMMA layers
__global__ void __launch_bounds__(256) cyl_dec0_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out) {
    constexpr int CI = 4, CO = 128, HS = 4, WS = 8, H0 = 8, W0 = 16;
    __shared__ float src[CI * HS * WS];
    __shared__ __align__(16) float wsm0[CO * CI * 9];
    __shared__ float up[CI][H0 + 2][W0 + 2];
    const int n = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < CI * HS * WS; i += 256) src[i] = g[(size_t)n * CI * HS * WS + i];
    for (int i = tid; i < CO * CI * 9; i += 256) wsm0[i] = __ldg(w + i);
    __syncthreads();
    const float sh = (float)(HS - 1) / (float)(H0 - 1), sw = (float)(WS - 1) / (float)(W0 - 1);
    for (int i = tid; i < CI * (H0 + 2) * (W0 + 2); i += 256) {
        const int xx = i % (W0 + 2), yy = (i / (W0 + 2)) % (H0 + 2), ci = i / ((W0 + 2) * (H0 + 2));
        const Lerp ly = lerp_idx(min(max(yy - 1, 0), H0 - 1), HS, sh), lx = lerp_idx(min(max(xx - 1, 0), W0 - 1), WS, sw);
        const float* ic = src + ci * HS * WS;
        up[ci][yy][xx] = ly.l0 * (lx.l0 * ic[ly.i0 * WS + lx.i0] + lx.l1 * ic[ly.i0 * WS + lx.i1]) +
                         ly.l1 * (lx.l0 * ic[ly.i1 * WS + lx.i0] + lx.l1 * ic[ly.i1 * WS + lx.i1]);
    }
    __syncthreads();
    const int px = tid & 127, half = tid >> 7;
    const int y = px >> 4, x = px & 15;
    float win[36];
#pragma unroll
    for (int ci = 0; ci < CI; ci++)
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) win[ci * 9 + dy * 3 + dx] = up[ci][y + dy][x + dx];
    for (int cq = half * 16; cq < half * 16 + 16; cq++) {
        float r4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int co = 4 * cq + k;
            const float4* wr = reinterpret_cast<const float4*>(wsm0 + co * 36);
            float acc = __ldg(bias + co);
#pragma unroll
            for (int q4 = 0; q4 < 9; q4++) {
                const float4 wv = wr[q4];
                acc = fmaf(win[4 * q4 + 0], wv.x, acc);
                acc = fmaf(win[4 * q4 + 1], wv.y, acc);
                acc = fmaf(win[4 * q4 + 2], wv.z, acc);
                acc = fmaf(win[4 * q4 + 3], wv.w, acc);
            }
            r4[k] = fmaxf(acc, 0.f);
            if (!CL4) out[((size_t)n * CO + co) * (H0 * W0) + px] = r4[k];
        }
        if (CL4)
            reinterpret_cast<float4*>(out)[((size_t)n * (CO / 4) + cq) * (H0 * W0) + px] = make_float4(r4[0], r4[1], r4[2], r4[3]);
    }
}

// ---- LayerNorm over the flattened encoder output (observableNetFC, embedding_cylinder.py:62-67) --------
__global__ void cyl_ln_kernel(const float* __restrict__ z, const float* __restrict__ lw, const float* __restrict__ lb,
                              float* __restrict__ g, int N, int E, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N) return;
    const float* zr = z + (size_t)warp * E;
    float s = 0.f;
    for (int c = lane; c < E; c += 32) s += zr[c];
    const float mean = warp_sum(s) / (float)E;
    float v = 0.f;
    for (int c = lane; c < E; c += 32) {
        const float d = zr[c] - mean;
        v = fmaf(d, d, v);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)E + eps);
    for (int c = lane; c < E; c += 32) g[(size_t)warp * E + c] = (zr[c] - mean) * rstd * __ldg(lw + c) + __ldg(lb + c);
}

// ---- K4 (cylinder): viscosity-conditioned banded Koopman operator --------------------------------------
// D = kMatrixDiagNet(100 visc) [E]; U = kMatrixUT(100 visc), L = kMatrixLT(100 visc) [502] with band o at
// offset sum_{q<o}(E-q); K[r][r+o] = U_o[r], K[r+o][r] = L_o[r]  (embedding_cylinder.py:95-101, 183-191).
struct KoopNets {
    const float *d0w, *d0b, *d2w, *d2b, *u0w, *u0b, *u2w, *u2b, *l0w, *l0b, *l2w, *l2b;
};

__global__ void cyl_koopman_kernel(KoopNets k, const float* __restrict__ g, const float* __restrict__ visc,
                                   float* __restrict__ gnext, float* __restrict__ kdiag, float* __restrict__ kmat, int E) {
    __shared__ float hd[50], hu[50], hl[50];
    extern __shared__ float band[];            // D[E] | U[nband] | L[nband]
    const int n = blockIdx.x, tid = threadIdx.x;
    const int nband = 4 * E - 10;
    const float v = 100.f * visc[n];
    if (tid < 50) hd[tid] = fmaxf(fmaf(__ldg(k.d0w + tid), v, __ldg(k.d0b + tid)), 0.f);
    else if (tid >= 64 && tid < 114) hu[tid - 64] = fmaxf(fmaf(__ldg(k.u0w + tid - 64), v, __ldg(k.u0b + tid - 64)), 0.f);
    else if (tid >= 128 && tid < 178) hl[tid - 128] = fmaxf(fmaf(__ldg(k.l0w + tid - 128), v, __ldg(k.l0b + tid - 128)), 0.f);
    __syncthreads();
    float* D = band;
    float* U = band + E;
    float* Lw = U + nband;
    for (int i = tid; i < E + 2 * nband; i += blockDim.x) {
        const float* wrow;
        const float* hh;
        float acc;
        if (i < E) { wrow = k.d2w + (size_t)i * 50; hh = hd; acc = 0.f; }
        else if (i < E + nband) { wrow = k.u2w + (size_t)(i - E) * 50; hh = hu; acc = 0.f; }
        else { wrow = k.l2w + (size_t)(i - E - nband) * 50; hh = hl; acc = 0.f; }
        for (int j = 0; j < 50; j++) acc = fmaf(hh[j], __ldg(wrow + j), acc);
        if (i < E) acc += __ldg(k.d2b + i);
        else if (i < E + nband) acc += __ldg(k.u2b + i - E);
        else acc += __ldg(k.l2b + i - E - nband);
        band[i] = acc;
    }
    __syncthreads();
    const float* gr = g + (size_t)n * E;
    for (int i = tid; i < E; i += blockDim.x) {
        float acc = 0.f;
        int off[NBAND + 1];
        off[1] = 0;
        for (int o = 2; o <= NBAND; o++) off[o] = off[o - 1] + (E - (o - 1));
        for (int o = NBAND; o >= 1; o--)
            if (i - o >= 0) acc = fmaf(Lw[off[o] + i - o], gr[i - o], acc);
        acc = fmaf(D[i], gr[i], acc);
        for (int o = 1; o <= NBAND; o++)
            if (i + o < E) acc = fmaf(U[off[o] + i], gr[i + o], acc);
        gnext[(size_t)n * E + i] = acc;
        if (kdiag) kdiag[(size_t)n * E + i] = D[i];
    }
    if (kmat) {
        float* Km = kmat + (size_t)n * E * E;
        for (int idx = tid; idx < E * E; idx += blockDim.x) {
            const int r = idx / E, c = idx % E;
            const int o = c - r;
            float val = 0.f;
            if (o == 0) val = D[r];
            else if (o > 0 && o <= NBAND) {
                int off = 0;
                for (int q = 1; q < o; q++) off += E - q;
                val = U[off + r];
            } else if (o < 0 && -o <= NBAND) {
                int off = 0;
                for (int q = 1; q < -o; q++) off += E - q;
                val = Lw[off + c];
            }
            Km[idx] = val;
        }
    }
}

int ensure_ws(trpx_cyl_t h, size_t bytes) {
    if (bytes <= h->ws_bytes) return TRPX_OK;
    if (h->ws) {
        TRPX_CUDA(cudaDeviceSynchronize());
        TRPX_CUDA(cudaFree(h->ws));
        h->ws = nullptr;
        h->ws_bytes = 0;
    }
    TRPX_CUDA(cudaMalloc(&h->ws, bytes));
    h->ws_bytes = bytes;
    return TRPX_OK;
}

template <int CT, bool UP>
int launch_tiled(const float* in, const float* w, const float* b, float* out, int N, int Cin, int Hs, int Ws, int Cout,
                 ConvEpilogue ep, cudaStream_t st) {
    const int H = UP ? 2 * Hs : Hs, W = UP ? 2 * Ws : Ws;
    dim3 grid(((H + TH - 1) / TH) * ((W + TW - 1) / TW), (Cout + 2 * CT - 1) / (2 * CT), N);
    conv3x3_tiled_kernel<CT, UP><<<grid, CONV_THREADS, 0, st>>>(in, w, b, out, Cin, Hs, Ws, Cout, ep);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

template <int NT>
int launch_up_mma(const float* in, const float* w, const float* b, float* out, int N, int Cin, int Hs, int Ws, int Cout,
                  cudaStream_t st) {
    const int H = 2 * Hs, W = 2 * Ws;
    const size_t smem = sizeof(float) * (2 * (size_t)MCC * MPLANE + 2 * (size_t)MCC * 9 * (NT * 8 + 8));
    TRPX_CUDA(cudaFuncSetAttribute(conv3x3_up_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(((H + TH - 1) / TH) * ((W + TW - 1) / TW), (Cout + NT * 8 - 1) / (NT * 8), N);
    conv3x3_up_mma_kernel<NT><<<grid, CONV_THREADS, smem, st>>>(in, w, b, out, Cin, Hs, Ws, Cout);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

template <int NT, bool UP>
int launch_pipe(trpx_cyl_t h, const float* in, const float* wimg, const float* b, float* out, int N, int Cin, int Hs,
                int Ws, int Cout, ConvEpilogue ep, cudaStream_t st) {
    const int H = UP ? 2 * Hs : Hs, W = UP ? 2 * Ws : Ws;
    const size_t plane = (size_t)N * Cin * (H + 2) * (W + 4);
    const int blocks = (int)std::min<size_t>((plane / 4 + 255) / 256, (size_t)148 * 16);
    cyl_upsplit_kernel<UP><<<blocks, 256, 0, st>>>(in, h->U, N * Cin, Hs, Ws, plane);
    TRPX_CUDA(cudaGetLastError());
    constexpr int COT = NT * 8, COP = COT + 8;
    const size_t smem = sizeof(float) * 2 * (2 * (size_t)MCC * MPLANE + 2 * (size_t)MCC * 9 * COP);
    TRPX_CUDA(cudaFuncSetAttribute(conv3x3_pipe_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((H / TH) * (W / TW), (Cout + COT - 1) / COT, N);
    conv3x3_pipe_mma_kernel<NT><<<grid, CONV_THREADS, smem, st>>>(h->U, plane, wimg, b, out, Cin, H, W, Cout, ep);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

template <bool UP>
int launch_naive(const float* in, const float* w, const float* b, float* out, int N, int Cin, int Hs, int Ws, int Cout,
                 int stride, int relu, cudaStream_t st) {
    const int Hi = UP ? 2 * Hs : Hs, Wi = UP ? 2 * Ws : Ws;
    const long long total = (long long)N * Cout * ((Hi - 1) / stride + 1) * ((Wi - 1) / stride + 1);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    conv3x3_naive_kernel<UP><<<blocks, 256, 0, st>>>(in, w, b, out, N, Cin, Hs, Ws, Cout, stride, relu);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

}  // namespace

extern "C" {

int trpx_cyl_create(int n_embd, float ln_eps, int device, trpx_cyl_t* out) {
    TRPX_REQUIRE(out != nullptr, "null argument");
    TRPX_REQUIRE(n_embd == 128, "n_embd=%d unsupported: the 64x128 conv stack flattens to 4*4*8=128", n_embd);
    int ndev = 0;
    TRPX_CUDA(cudaGetDeviceCount(&ndev));
    TRPX_REQUIRE(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
    DeviceGuard guard(device);
    trpx_cyl* h = new (std::nothrow) trpx_cyl();
    TRPX_REQUIRE(h != nullptr, "out of host memory");
    h->E = n_embd;
    h->eps = ln_eps;
    h->device = device;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    size_t n[36];
    tensor_sizes(n_embd, n);
    size_t o = 0;
    for (int i = 0; i < 36; i++) {
        h->off[i] = o;
        o += (n[i] + 3) & ~(size_t)3;          // keep every tensor 16-byte aligned
    }
    h->total = o;
    // cylinder mask, embedding_cylinder.py:38-39: sqrt(X^2 + Y^2) < 1 on linspace(-2,14,128) x linspace(-4,4,64)
    std::vector<unsigned char> mk(CYL_H * CYL_W);
    for (int y = 0; y < CYL_H; y++)
        for (int x = 0; x < CYL_W; x++) {
            const double X = (x == CYL_W - 1) ? 14.0 : -2.0 + x * (16.0 / (CYL_W - 1));
            const double Y = (y == CYL_H - 1) ? 4.0 : -4.0 + y * (8.0 / (CYL_H - 1));
            mk[y * CYL_W + x] = std::sqrt(X * X + Y * Y) < 1.0 ? 1 : 0;
        }
    if (cudaMalloc(&h->blob, sizeof(float) * h->total) != cudaSuccess ||
        cudaMalloc(&h->mask, mk.size()) != cudaSuccess ||
        cudaMemcpy(h->mask, mk.data(), mk.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(h->blob);
        cudaFree(h->mask);
        delete h;
        return set_error(TRPX_E_CUDA, "allocation of the cylinder handle failed");
    }
    *out = h;
    return TRPX_OK;
}

int trpx_cyl_destroy(trpx_cyl_t h) {
    if (!valid(h)) return set_error(TRPX_E_STATE, "trpx_cyl_destroy: invalid handle");
    DeviceGuard guard(h->device);
    cudaFree(h->blob);
    cudaFree(h->mask);
    cudaFree(h->ws);
    cudaFree(h->wimg);
    cudaFree(h->U);
    cudaFree(h->uimg);
    cudaFree(h->eimg);
    cudaFree(h->tws);
    h->magic = 0;
    delete h;
    return TRPX_OK;
}

int trpx_cyl_load_weights(trpx_cyl_t h, const float* const* w, int n, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    TRPX_REQUIRE(w != nullptr && n == 36, "expected 36 weight tensors, got %d", n);
    for (int i = 0; i < n; i++) TRPX_REQUIRE(w[i] != nullptr, "weight tensor %d is null", i);
    DeviceGuard guard(h->device);
    size_t sz[36];
    tensor_sizes(h->E, sz);
    for (int i = 0; i < 36; i++)
        TRPX_CUDA(cudaMemcpyAsync(h->blob + h->off[i], w[i], sz[i] * sizeof(float), cudaMemcpyDeviceToDevice,
                                  (cudaStream_t)stream));
    {   // weight images of the three heavy decoder layers for the pipelined tensor-core path
        const int cin[4] = {128, 64, 32, 16}, cout[4] = {64, 32, 16, 3}, cot[4] = {32, 32, 16, 8};
        size_t tot = 0;
        for (int i = 0; i < 4; i++) {
            h->wimg_off[i] = tot;
            tot += (size_t)((cout[i] + cot[i] - 1) / cot[i]) * (cin[i] / MCC) * 2 * MCC * 9 * (cot[i] + 8);
        }
        if (h->wimg == nullptr) TRPX_CUDA(cudaMalloc(&h->wimg, tot * sizeof(float)));
        for (int i = 0; i < 4; i++) {
            cyl_wimage_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(h->blob + h->off[T_DEC + 2 + 2 * i],
                                                                  h->wimg + h->wimg_off[i], cin[i], cout[i], cot[i]);
            TRPX_CUDA(cudaGetLastError());
        }
    }
    {   // weight images of the tcgen05 layers
        size_t tot = 0;
        for (int i = 0; i < UMMA_LAYERS; i++) { h->uimg_off[i] = tot; tot += umma_wimg_floats(i); }
        h->uimg_off[UMMA_LAYERS] = tot;
        if (h->uimg == nullptr) TRPX_CUDA(cudaMalloc(&h->uimg, tot * sizeof(float)));
        for (int i = 0; i < UMMA_LAYERS; i++) {
            int rc = umma_build_wimg(i, h->blob + h->off[T_DEC + 2 + 2 * i], h->uimg + h->uimg_off[i], (cudaStream_t)stream);
            if (rc) return rc;
        }
    }
    {   // ... and of the encoder layers
        size_t tot = 0;
        for (int i = 0; i < UMMA_ENC_LAYERS; i++) { h->eimg_off[i] = tot; tot += umma_enc_wimg_floats(i); }
        h->eimg_off[UMMA_ENC_LAYERS] = tot;
        if (h->eimg == nullptr) TRPX_CUDA(cudaMalloc(&h->eimg, tot * sizeof(float)));
        for (int i = 0; i < UMMA_ENC_LAYERS; i++) {
            int rc = umma_enc_build_wimg(i, h->blob + h->off[T_ENC + 2 * i], h->eimg + h->eimg_off[i], (cudaStream_t)stream);
            if (rc) return rc;
        }
    }
    h->loaded = true;
    return TRPX_OK;
}

int trpx_cyl_set_conv_mode(trpx_cyl_t h, int mode) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    TRPX_REQUIRE(mode >= 0 && mode <= 3, "conv mode must be 0 (3xTF32 in-kernel build), 1 (FP32 SIMT), 2 (pipelined 3xTF32 mma.sync) or 3 (tcgen05)");
    h->conv_mode = mode;
    return TRPX_OK;
}

int trpx_cyl_set_chunk(trpx_cyl_t h, int frames) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    TRPX_REQUIRE(frames >= 1 && frames <= 65535, "chunk must be in [1, 65535]");
    h->chunk = frames;
    h->chunk_user = true;
    return TRPX_OK;
}

int trpx_cyl_embed(trpx_cyl_t h, const float* x, const float* visc, float* g, int N, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "weights not loaded");
    TRPX_REQUIRE(N >= 0 && (N == 0 || (x && visc && g)), "bad arguments");
    if (N == 0) return TRPX_OK;
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = CYL_H * CYL_W;
    const int c4 = h->E / 32;
    // per-frame activation sizes: normalised input, 4 strided convs, final conv
    const size_t a0 = 4 * HW, a1 = 16 * HW / 4, a2 = 32 * HW / 16, a3 = 64 * HW / 64, a4 = 128 * HW / 256, a5 = (size_t)c4 * 32;
    const size_t per = a0 + a1 + a2 + a3 + a4 + a5;
    const float* B = h->blob;
    if (h->conv_mode == 3) {
        // tcgen05 implicit GEMMs: the stride-2 layers in space-to-depth form, each epilogue writing the next layer's image
        const int chunk3 = std::min(N, h->chunk_user ? h->chunk : 1024);
        int rc3 = ensure_ws(h, (umma_enc_ws_floats(chunk3)) * sizeof(float));
        if (rc3) return rc3;
        const float* wi[UMMA_ENC_LAYERS];
        const float* bi[UMMA_ENC_LAYERS];
        for (int i = 0; i < UMMA_ENC_LAYERS; i++) { wi[i] = h->eimg + h->eimg_off[i]; bi[i] = B + h->off[T_ENC + 2 * i + 1]; }
        for (int n0 = 0; n0 < N; n0 += chunk3) {
            const int nn = std::min(chunk3, N - n0);
            float* z = h->ws + umma_enc_ws_floats(chunk3) - (size_t)chunk3 * 128;
            if ((rc3 = umma_encoder(x + (size_t)n0 * 3 * HW, visc + n0, B + h->off[T_MU], B + h->off[T_STD], wi, bi, h->ws, z, nn,
                                    h->sm_count, st))) return rc3;
            cyl_ln_kernel<<<(nn * 32 + 255) / 256, 256, 0, st>>>(z, B + h->off[T_LNW], B + h->off[T_LNB], g + (size_t)n0 * h->E, nn, h->E, h->eps);
            TRPX_CUDA(cudaGetLastError());
        }
        return TRPX_OK;
    }
    const int chunk = std::min(N, h->chunk);
    int rc = ensure_ws(h, per * chunk * sizeof(float));
    if (rc) return rc;
    for (int n0 = 0; n0 < N; n0 += chunk) {
        const int nn = std::min(chunk, N - n0);
        float* p0 = h->ws;
        float* p1 = p0 + a0 * chunk;
        float* p2 = p1 + a1 * chunk;
        float* p3 = p2 + a2 * chunk;
        float* p4 = p3 + a3 * chunk;
        float* p5 = p4 + a4 * chunk;
        cyl_prep_kernel<<<std::min(nn * 32, 148 * 8), 256, 0, st>>>(x + (size_t)n0 * 3 * HW, visc + n0, B + h->off[T_MU],
                                                                    B + h->off[T_STD], p0, nn);
        TRPX_CUDA(cudaGetLastError());
        if ((rc = launch_naive<false>(p0, B + h->off[T_ENC + 0], B + h->off[T_ENC + 1], p1, nn, 4, 64, 128, 16, 2, 1, st))) return rc;
        if ((rc = launch_naive<false>(p1, B + h->off[T_ENC + 2], B + h->off[T_ENC + 3], p2, nn, 16, 32, 64, 32, 2, 1, st))) return rc;
        if ((rc = launch_naive<false>(p2, B + h->off[T_ENC + 4], B + h->off[T_ENC + 5], p3, nn, 32, 16, 32, 64, 2, 1, st))) return rc;
        if ((rc = launch_naive<false>(p3, B + h->off[T_ENC + 6], B + h->off[T_ENC + 7], p4, nn, 64, 8, 16, 128, 2, 1, st))) return rc;
        if ((rc = launch_naive<false>(p4, B + h->off[T_ENC + 8], B + h->off[T_ENC + 9], p5, nn, 128, 4, 8, c4, 1, 0, st))) return rc;
        cyl_ln_kernel<<<(nn * 32 + 255) / 256, 256, 0, st>>>(p5, B + h->off[T_LNW], B + h->off[T_LNB], g + (size_t)n0 * h->E,
                                                             nn, h->E, h->eps);
        TRPX_CUDA(cudaGetLastError());
    }
    return TRPX_OK;
}

int trpx_cyl_recover(trpx_cyl_t h, const float* g, float* x, int N, int apply_mask, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "weights not loaded");
    TRPX_REQUIRE(N >= 0 && (N == 0 || (x && g)), "bad arguments");
    if (N == 0) return TRPX_OK;
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = CYL_H * CYL_W;
    const int c4 = h->E / 32;
    const size_t a1 = 128 * 8 * 16, a2 = 64 * 16 * 32, a3 = 32 * 32 * 64, a4 = 16 * 64 * 128;
    const size_t per = a1 + a2 + a3 + a4;
    // frames per pass: 64 keeps one pass's intermediates L2-resident for the SIMT path; the pipelined path is
    // tensor-bound and prefers fewer, larger launches (measured: 256)
    const int chunk = std::min(N, (h->conv_mode >= 2 && !h->chunk_user) ? 256 : h->chunk);
    int rc = ensure_ws(h, per * chunk * sizeof(float));
    if (rc) return rc;
    if (h->conv_mode == 2) {
        // largest layer input after upsampling: [32][66][132] floats per frame, two planes
        const size_t need = (size_t)chunk * 2 * 32 * 66 * 132 * sizeof(float);
        if (need > h->U_bytes) {
            if (h->U) { TRPX_CUDA(cudaDeviceSynchronize()); TRPX_CUDA(cudaFree(h->U)); h->U = nullptr; h->U_bytes = 0; }
            TRPX_CUDA(cudaMalloc(&h->U, need));
            h->U_bytes = need;
        }
    }
    const float* B = h->blob;
    ConvEpilogue relu{1, 0, nullptr, nullptr, nullptr};
    ConvEpilogue fin{0, 1, B + h->off[T_MU], B + h->off[T_STD], apply_mask ? h->mask : nullptr};
    for (int n0 = 0; n0 < N; n0 += chunk) {
        const int nn = std::min(chunk, N - n0);
        float* p1 = h->ws;
        float* p2 = p1 + a1 * chunk;
        float* p3 = p2 + a2 * chunk;
        float* p4 = p3 + a3 * chunk;
        if (h->conv_mode == 3) {       // tcgen05 implicit GEMMs (csrc/cyl_umma.cu); activations channels-last-4
            cyl_dec0_kernel<true><<<nn, 256, 0, st>>>(g + (size_t)n0 * h->E, B + h->off[T_DEC + 0], B + h->off[T_DEC + 1], p1);
            TRPX_CUDA(cudaGetLastError());
            float* bufs[5] = {p1, p2, p3, p4, x + (size_t)n0 * 3 * HW};
            for (int l = 0; l < UMMA_LAYERS; l++) {
                rc = umma_conv_layer(l, bufs[l], h->uimg + h->uimg_off[l], B + h->off[T_DEC + 3 + 2 * l], bufs[l + 1], nn,
                                     B + h->off[T_MU], B + h->off[T_STD], apply_mask ? h->mask : nullptr, h->sm_count, st);
                if (rc) return rc;
            }
            continue;
        }
        if (h->conv_mode == 2) {       // pipelined 3xTF32 tensor-core path, all five layers
            cyl_dec0_kernel<false><<<nn, 256, 0, st>>>(g + (size_t)n0 * h->E, B + h->off[T_DEC + 0], B + h->off[T_DEC + 1], p1);
            TRPX_CUDA(cudaGetLastError());
            if ((rc = launch_pipe<4, true>(h, p1, h->wimg + h->wimg_off[0], B + h->off[T_DEC + 3], p2, nn, 128, 8, 16, 64, relu, st))) return rc;
            if ((rc = launch_pipe<4, true>(h, p2, h->wimg + h->wimg_off[1], B + h->off[T_DEC + 5], p3, nn, 64, 16, 32, 32, relu, st))) return rc;
            if ((rc = launch_pipe<2, true>(h, p3, h->wimg + h->wimg_off[2], B + h->off[T_DEC + 7], p4, nn, 32, 32, 64, 16, relu, st))) return rc;
            // final 16 -> 3 conv: the FP32 SIMT tile kernel is faster than the pipeline here (172 vs 216 us per 256
            // frames: with 3 output channels the layer is bound by moving its input, not by math)
            if (h->final_mma) {
                if ((rc = launch_pipe<1, false>(h, p4, h->wimg + h->wimg_off[3], B + h->off[T_DEC + 9], x + (size_t)n0 * 3 * HW, nn, 16, 64, 128, 3, fin, st))) return rc;
            } else {
                if ((rc = launch_tiled<2, false>(p4, B + h->off[T_DEC + 8], B + h->off[T_DEC + 9], x + (size_t)n0 * 3 * HW, nn, 16, 64, 128, 3, fin, st))) return rc;
            }
            continue;
        }
        // g.view(-1, 4, 4, 8) -> up x2 -> conv(4->128) -> relu        (embedding_cylinder.py:71-73)
        if ((rc = launch_naive<true>(g + (size_t)n0 * h->E, B + h->off[T_DEC + 0], B + h->off[T_DEC + 1], p1, nn, c4, 4, 8, 128, 1, 1, st))) return rc;
        if (h->conv_mode == 1) {       // FP32 SIMT direct convolution (default)
            if ((rc = launch_tiled<8, true>(p1, B + h->off[T_DEC + 2], B + h->off[T_DEC + 3], p2, nn, 128, 8, 16, 64, relu, st))) return rc;
            if ((rc = launch_tiled<8, true>(p2, B + h->off[T_DEC + 4], B + h->off[T_DEC + 5], p3, nn, 64, 16, 32, 32, relu, st))) return rc;
            if ((rc = launch_tiled<8, true>(p3, B + h->off[T_DEC + 6], B + h->off[T_DEC + 7], p4, nn, 32, 32, 64, 16, relu, st))) return rc;
        } else {                       // 3xTF32 tensor-core implicit GEMM (opt-in: measured equal speed, 8x the error)
            if ((rc = launch_up_mma<8>(p1, B + h->off[T_DEC + 2], B + h->off[T_DEC + 3], p2, nn, 128, 8, 16, 64, st))) return rc;
            if ((rc = launch_up_mma<4>(p2, B + h->off[T_DEC + 4], B + h->off[T_DEC + 5], p3, nn, 64, 16, 32, 32, st))) return rc;
            if ((rc = launch_up_mma<2>(p3, B + h->off[T_DEC + 6], B + h->off[T_DEC + 7], p4, nn, 32, 32, 64, 16, st))) return rc;
        }
        if ((rc = launch_tiled<2, false>(p4, B + h->off[T_DEC + 8], B + h->off[T_DEC + 9], x + (size_t)n0 * 3 * HW, nn, 16, 64, 128, 3, fin, st))) return rc;
    }
    return TRPX_OK;
}

int trpx_cyl_koopman(trpx_cyl_t h, const float* g, const float* visc, float* gnext, float* kdiag, float* kmat, int N,
                     void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "weights not loaded");
    TRPX_REQUIRE(N >= 0 && (N == 0 || (g && visc && gnext)), "bad arguments");
    TRPX_REQUIRE(g != gnext, "in-place Koopman step is not supported");
    if (N == 0) return TRPX_OK;
    DeviceGuard guard(h->device);
    const float* B = h->blob;
    KoopNets k{B + h->off[T_KD], B + h->off[T_KD + 1], B + h->off[T_KD + 2], B + h->off[T_KD + 3],
               B + h->off[T_KU], B + h->off[T_KU + 1], B + h->off[T_KU + 2], B + h->off[T_KU + 3],
               B + h->off[T_KL], B + h->off[T_KL + 1], B + h->off[T_KL + 2], B + h->off[T_KL + 3]};
    const int nband = 4 * h->E - 10;
    const size_t smem = sizeof(float) * (h->E + 2 * nband);
    cyl_koopman_kernel<<<N, 256, smem, (cudaStream_t)stream>>>(k, g, visc, gnext, kdiag, kmat, h->E);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

}  // extern "C"
