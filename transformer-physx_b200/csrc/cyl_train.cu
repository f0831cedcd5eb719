// Cylinder Koopman embedding TRAINING step: loss and gradient of CylinderEmbeddingTrainer.forward
// (trphysx/embedding/embedding_cylinder.py:236-281) for every parameter of CylinderEmbedding, in one native call.
//
//   loss = w_rec MSE(x_0, D(E(x_0))) + sum_{t>=1} [ w_dyn MSE(Dm(K g_{t-1}), x_t) + w_rec MSE(D(E(x_t)), x_t) + w_reg sum(K^2) ]
//   g_0 = E(x_0, visc), g_t = K(visc) g_{t-1};  D = recoveryNet + un-normalise WITHOUT the cylinder mask (the reference's
//   forward(), fact 8), Dm = recover() WITH it;  K [B,128,128] banded, from the three viscosity nets (:172-196);
//   every MSE is a mean over B*3*64*128 values (nn.MSELoss); the regulariser sums over the batch (:272).
//
// Everything is FP32 SIMT and deterministic (no atomics: weight and bias gradients are split-K partial sums added in
// a fixed order).  Forward activations are kept (planar [n][c][h][w]); the backward of a layer is three kernels:
// weight gradient (tiled, the bilinear x2 upsample of the layer input is recomputed on the fly), data gradient at the
// conv-input resolution (gather form, replicate padding folded into the border pixels), adjoint of the bilinear upsample
// (gather form).  This is the first, correctness-first version of SURVEY.md 8f rank 3 (no tensor cores yet).
#include "cyl.cuh"
#include <algorithm>
#include <vector>

using namespace trpx;

namespace {

constexpr int HW = CYL_H * CYL_W;

// ------------------------------------------------------------------------------------------------------------------
// forward pieces
// ------------------------------------------------------------------------------------------------------------------
// states [NF][3][64][128] + visc[b] -> normalised 4-channel input (embedding_cylinder.py:149-150, 198-200)
__global__ void tr_prep_kernel(const float* __restrict__ x, const float* __restrict__ visc, const float* __restrict__ mu,
                               const float* __restrict__ sd, float* __restrict__ out, int NF, int T) {
    const long long total = (long long)NF * 4 * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int pix = (int)(i % HW);
        const int c = (int)((i / HW) % 4);
        const int n = (int)(i / (4LL * HW));
        const float v = (c < 3) ? x[((long long)n * 3 + c) * HW + pix] : visc[n / T];
        out[i] = (v - __ldg(mu + c)) / __ldg(sd + c);
    }
}

// direct 3x3 conv, replicate padding, optional bilinear x2 upsample of the input, one thread per output element
template <bool UP>
__global__ void tr_conv_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                   float* __restrict__ out, int N, int Cin, int Hs, int Ws, int Cout, int stride, int relu) {
    const int Hi = UP ? 2 * Hs : Hs, Wi = UP ? 2 * Ws : Ws;
    const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
    const long long total = (long long)N * Cout * Ho * Wo;
    const float sh = UP ? (float)(Hs - 1) / (float)(Hi - 1) : 0.f, sw = UP ? (float)(Ws - 1) / (float)(Wi - 1) : 0.f;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int xo = (int)(idx % Wo), yo = (int)((idx / Wo) % Ho);
        const int co = (int)((idx / ((long long)Wo * Ho)) % Cout), n = (int)(idx / ((long long)Wo * Ho * Cout));
        float acc = __ldg(bias + co);
        const float* wp = w + (size_t)co * Cin * 9;
        const float* ip = in + (size_t)n * Cin * Hs * Ws;
        Lerp ly[3], lx[3];
        int yy[3], xx[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            yy[d] = min(max(yo * stride + d - 1, 0), Hi - 1);
            xx[d] = min(max(xo * stride + d - 1, 0), Wi - 1);
            if (UP) { ly[d] = lerp_idx(yy[d], Hs, sh); lx[d] = lerp_idx(xx[d], Ws, sw); }
        }
        for (int ci = 0; ci < Cin; ci++) {
            const float* ic = ip + (size_t)ci * Hs * Ws;
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    float v;
                    if (UP) {
                        const Lerp a = ly[dy], b = lx[dx];
                        v = a.l0 * (b.l0 * ic[a.i0 * Ws + b.i0] + b.l1 * ic[a.i0 * Ws + b.i1]) +
                            a.l1 * (b.l0 * ic[a.i1 * Ws + b.i0] + b.l1 * ic[a.i1 * Ws + b.i1]);
                    } else {
                        v = ic[yy[dy] * Ws + xx[dx]];
                    }
                    acc = fmaf(v, __ldg(wp + ci * 9 + dy * 3 + dx), acc);
                }
        }
        out[idx] = relu ? fmaxf(acc, 0.f) : acc;
    }
}

// LayerNorm over the 128 flattened encoder outputs; keeps mean / rstd for the backward
__global__ void tr_ln_fwd_kernel(const float* __restrict__ z, const float* __restrict__ lw, const float* __restrict__ lb,
                                 float* __restrict__ g, float* __restrict__ stat, int N, int E, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N) return;
    const float* zr = z + (size_t)warp * E;
    float s = 0.f;
    for (int c = lane; c < E; c += 32) s += zr[c];
    const float mean = warp_sum(s) / (float)E;
    float v = 0.f;
    for (int c = lane; c < E; c += 32) { const float d = zr[c] - mean; v = fmaf(d, d, v); }
    const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)E + eps);
    for (int c = lane; c < E; c += 32) g[(size_t)warp * E + c] = (zr[c] - mean) * rstd * __ldg(lw + c) + __ldg(lb + c);
    if (lane == 0) { stat[2 * warp] = mean; stat[2 * warp + 1] = rstd; }
}

struct KoopNets {
    const float *w0[3], *b0[3], *w2[3], *b2[3];      // 0: diagonal net, 1: upper bands, 2: lower bands
};

// Per sample: hidden activations of the three viscosity nets, the bands D | U | L, sum(K^2), and the Koopman chain
// g_t = K g_{t-1}, t = 1..T-1 (embedding_cylinder.py:182-193).  One CTA per sample.
__global__ void tr_koop_fwd_kernel(KoopNets k, const float* __restrict__ visc, const float* __restrict__ G, float* __restrict__ hid,
                                   float* __restrict__ bands, float* __restrict__ ksq, float* __restrict__ gpred, int B, int T, int E) {
    extern __shared__ float sm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nband = 4 * E - 10, nb = E + 2 * nband;
    float* h = sm;                  // [3][50]
    float* band = sm + 150;         // [nb]
    float* ga = band + nb;          // [E]
    float* gb = ga + E;             // [E]
    float* red = gb + E;            // [blockDim/32]
    const float v = 100.f * visc[b];
    if (tid < 150) {
        const int net = tid / 50, j = tid % 50;
        const float a = fmaxf(fmaf(__ldg(k.w0[net] + j), v, __ldg(k.b0[net] + j)), 0.f);
        h[tid] = a;
        hid[(size_t)b * 150 + tid] = a;
    }
    __syncthreads();
    float sq = 0.f;
    for (int i = tid; i < nb; i += blockDim.x) {
        const int net = i < E ? 0 : (i < E + nband ? 1 : 2);
        const int r = i < E ? i : (i < E + nband ? i - E : i - E - nband);
        const float* wrow = k.w2[net] + (size_t)r * 50;
        float acc = 0.f;
        for (int j = 0; j < 50; j++) acc = fmaf(h[net * 50 + j], __ldg(wrow + j), acc);
        acc += __ldg(k.b2[net] + r);
        band[i] = acc;
        bands[(size_t)b * nb + i] = acc;
        sq = fmaf(acc, acc, sq);
    }
    sq = warp_sum(sq);
    if ((tid & 31) == 0) red[tid >> 5] = sq;
    for (int i = tid; i < E; i += blockDim.x) ga[i] = G[(size_t)(b * T) * E + i];       // g_0 = E(x_0) of this sample
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += red[i];
        ksq[b] = s;
    }
    const float* D = band;
    const float* U = band + E;
    const float* Lw = U + nband;
    int off[NBAND + 1];
    off[1] = 0;
    for (int o = 2; o <= NBAND; o++) off[o] = off[o - 1] + (E - (o - 1));
    float* cur = ga;
    float* nxt = gb;
    for (int t = 1; t < T; t++) {
        for (int i = tid; i < E; i += blockDim.x) {
            float acc = 0.f;
            for (int o = NBAND; o >= 1; o--)
                if (i - o >= 0) acc = fmaf(Lw[off[o] + i - o], cur[i - o], acc);
            acc = fmaf(D[i], cur[i], acc);
            for (int o = 1; o <= NBAND; o++)
                if (i + o < E) acc = fmaf(U[off[o] + i], cur[i + o], acc);
            nxt[i] = acc;
            gpred[((size_t)(t - 1) * B + b) * E + i] = acc;
        }
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// loss and its gradient at the decoder output (normalised space): xhat = sd*y + mu (masked decodes: 0 in the cylinder)
// ------------------------------------------------------------------------------------------------------------------
__global__ void tr_loss_grad_kernel(const float* __restrict__ y5, const float* __restrict__ states, const float* __restrict__ mu,
                                    const float* __restrict__ sd, const unsigned char* __restrict__ mask, float* __restrict__ dy5,
                                    float* __restrict__ part, int NF, int B, int T, float c_rec, float c_dyn) {
    // grid: (blocks per decode, ND); part[(d * gridDim.x + blockIdx.x)] = sum of squared errors of this block
    const int d = blockIdx.y;
    const bool dyn = d >= NF;
    int frame = d;
    if (dyn) { const int j = d - NF; frame = (j % B) * T + 1 + j / B; }
    const float coef = dyn ? c_dyn : c_rec;
    float s = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * HW; i += gridDim.x * blockDim.x) {
        const int c = i / HW, pix = i - c * HW;
        float grad = 0.f;
        float xh = __ldg(sd + c) * y5[(size_t)d * 3 * HW + i] + __ldg(mu + c);
        const bool m = dyn && mask[pix] != 0;
        if (m) xh = 0.f;
        const float diff = xh - states[(size_t)frame * 3 * HW + i];
        s = fmaf(diff, diff, s);
        if (!m) grad = coef * 2.f * diff * __ldg(sd + c);
        dy5[(size_t)d * 3 * HW + i] = grad;
    }
    __shared__ float red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[i];
        part[(size_t)d * gridDim.x + blockIdx.x] = t;
    }
}

// loss2[0] = w_rec*rec + w_dyn*dyn + w_reg*(T-1)*sum_b ksq[b];  loss2[1] = rec   (rec, dyn: sums of per-step MSEs)
__global__ void tr_loss_final_kernel(const float* __restrict__ part, int per, int NF, int ND, const float* __restrict__ ksq, int B,
                                     int T, float w_rec, float w_dyn, float w_reg, float* __restrict__ loss2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double rec = 0.0, dyn = 0.0, kk = 0.0;
    for (int d = 0; d < ND; d++) {
        double s = 0.0;
        for (int i = 0; i < per; i++) s += (double)part[(size_t)d * per + i];
        if (d < NF) rec += s; else dyn += s;
    }
    for (int b = 0; b < B; b++) kk += (double)ksq[b];
    const double norm = 1.0 / ((double)B * 3.0 * HW);
    rec *= norm; dyn *= norm;
    loss2[0] = (float)((double)w_rec * rec + (double)w_dyn * dyn + (double)w_reg * (double)(T - 1) * kk);
    loss2[1] = (float)rec;
}

// ------------------------------------------------------------------------------------------------------------------
// backward pieces
// ------------------------------------------------------------------------------------------------------------------
__global__ void tr_relu_bwd_kernel(float* __restrict__ g, const float* __restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (!(y[i] > 0.f)) g[i] = 0.f;
}

// Data gradient of a 3x3 conv (stride S, replicate padding) at the conv-input resolution Hi x Wi: gather form, the
// out-of-range taps of a border pixel are folded into it.  dz [N][Cout][Ho][Wo] -> dx [N][Cin][Hi][Wi].
template <int S>
__global__ void tr_conv_bwd_data_kernel(const float* __restrict__ dz, const float* __restrict__ w, float* __restrict__ dx, int N,
                                        int Cin, int Hi, int Wi, int Cout, int Ho, int Wo) {
    const long long total = (long long)N * Cin * Hi * Wi;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int xi = (int)(idx % Wi), yi = (int)((idx / Wi) % Hi);
        const int ci = (int)((idx / ((long long)Wi * Hi)) % Cin), n = (int)(idx / ((long long)Wi * Hi * Cin));
        // (yo, dy) pairs whose clamped input row is yi, same for columns
        int ys[6], dys[6], ny = 0, xs[6], dxs[6], nx = 0;
#pragma unroll
        for (int e = 0; e < 3; e++) {
            const int uy = e == 0 ? yi : (e == 1 ? -1 : Hi);            // unclamped rows that read row yi
            if ((e == 1 && yi != 0) || (e == 2 && yi != Hi - 1)) continue;
#pragma unroll
            for (int dy = 0; dy < 3; dy++) {
                const int ty = uy - dy + 1;
                if (ty < 0 || (ty % S) != 0) continue;
                const int yo = ty / S;
                if (yo >= Ho) continue;
                ys[ny] = yo; dys[ny] = dy; ny++;
            }
        }
#pragma unroll
        for (int e = 0; e < 3; e++) {
            const int ux = e == 0 ? xi : (e == 1 ? -1 : Wi);
            if ((e == 1 && xi != 0) || (e == 2 && xi != Wi - 1)) continue;
#pragma unroll
            for (int dxx = 0; dxx < 3; dxx++) {
                const int tx = ux - dxx + 1;
                if (tx < 0 || (tx % S) != 0) continue;
                const int xo = tx / S;
                if (xo >= Wo) continue;
                xs[nx] = xo; dxs[nx] = dxx; nx++;
            }
        }
        float acc = 0.f;
        const float* zp = dz + (size_t)n * Cout * Ho * Wo;
        for (int co = 0; co < Cout; co++) {
            const float* zc = zp + (size_t)co * Ho * Wo;
            const float* wc = w + ((size_t)co * Cin + ci) * 9;
            for (int a = 0; a < ny; a++)
                for (int b = 0; b < nx; b++) acc = fmaf(zc[ys[a] * Wo + xs[b]], __ldg(wc + dys[a] * 3 + dxs[b]), acc);
        }
        dx[idx] = acc;
    }
}

// adjoint of the bilinear x2 upsample (align_corners=True): dU [NC][2Hs][2Ws] -> dx [NC][Hs][Ws], gather form
__global__ void tr_upsample_bwd_kernel(const float* __restrict__ dU, float* __restrict__ dx, long long NC, int Hs, int Ws) {
    const int H = 2 * Hs, W = 2 * Ws;
    const float sh = (float)(Hs - 1) / (float)(H - 1), sw = (float)(Ws - 1) / (float)(W - 1);
    const long long total = NC * Hs * Ws;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % Ws), i = (int)((idx / Ws) % Hs);
        const long long nc = idx / ((long long)Ws * Hs);
        const float* up = dU + nc * H * W;
        float wy[8], wx[8];
        int y0 = max(2 * i - 3, 0), y1 = min(2 * i + 4, H - 1), x0 = max(2 * j - 3, 0), x1 = min(2 * j + 4, W - 1);
        for (int y = y0; y <= y1; y++) {
            const Lerp l = lerp_idx(y, Hs, sh);
            wy[y - y0] = (l.i0 == i ? l.l0 : 0.f) + (l.i1 == i ? l.l1 : 0.f);
        }
        for (int x = x0; x <= x1; x++) {
            const Lerp l = lerp_idx(x, Ws, sw);
            wx[x - x0] = (l.i0 == j ? l.l0 : 0.f) + (l.i1 == j ? l.l1 : 0.f);
        }
        float acc = 0.f;
        for (int y = y0; y <= y1; y++) {
            if (wy[y - y0] == 0.f) continue;
            float r = 0.f;
            for (int x = x0; x <= x1; x++) r = fmaf(wx[x - x0], up[y * W + x], r);
            acc = fmaf(wy[y - y0], r, acc);
        }
        dx[idx] = acc;
    }
}

// Weight gradient of a 3x3 conv: dW[co][ci][dy][dx] = sum_{n,yo,xo} dz[n][co][yo][xo] * xpad[n][ci][yo*S+dy-1][xo*S+dx-1],
// xpad = replicate-padded (and, if UP, bilinearly x2 upsampled) layer input.  CTA tile: 32 output x 8 input channels,
// thread = one (co, ci) pair with its 9 taps in registers; grid.z slices the (n, yo) rows (split-K); partial sums go to
// part[z][Cout][Cin][9] and are added in a fixed order by tr_reduce_kernel.
constexpr int BW_CO = 32, BW_CI = 8, BW_MAXW = 128;
template <int S, bool UP>
__global__ void __launch_bounds__(256) tr_conv_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                                 float* __restrict__ part, int N, int Cin, int Hs, int Ws, int Cout,
                                                                 int Ho, int Wo) {
    __shared__ float dzs[BW_CO][BW_MAXW + 1];
    __shared__ float xs[BW_CI][3][S * BW_MAXW + 3];
    const int Hi = UP ? 2 * Hs : Hs, Wi = UP ? 2 * Ws : Ws;
    const float sh = UP ? (float)(Hs - 1) / (float)(Hi - 1) : 0.f, sw = UP ? (float)(Ws - 1) / (float)(Wi - 1) : 0.f;
    const int co0 = blockIdx.x * BW_CO, ci0 = blockIdx.y * BW_CI;
    const int tid = threadIdx.x, tco = tid >> 3, tci = tid & 7;
    const int wx = S * (Wo - 1) + 3;                       // padded input columns feeding one output row
    float acc[9];
#pragma unroll
    for (int i = 0; i < 9; i++) acc[i] = 0.f;
    const int rows = N * Ho;
    for (int r = blockIdx.z; r < rows; r += gridDim.z) {
        const int n = r / Ho, yo = r - n * Ho;
        for (int i = tid; i < BW_CO * Wo; i += 256) {
            const int c = i / Wo, xo = i - c * Wo;
            dzs[c][xo] = (co0 + c < Cout) ? dz[(((size_t)n * Cout + co0 + c) * Ho + yo) * Wo + xo] : 0.f;
        }
        for (int i = tid; i < BW_CI * 3 * wx; i += 256) {
            const int xx = i % wx, rr = i / wx, dy = rr % 3, c = rr / 3;
            float v = 0.f;
            if (ci0 + c < Cin) {
                const int yy = min(max(yo * S + dy - 1, 0), Hi - 1), xc = min(max(xx - 1, 0), Wi - 1);
                const float* ic = x + ((size_t)n * Cin + ci0 + c) * Hs * Ws;
                if (UP) {
                    const Lerp a = lerp_idx(yy, Hs, sh), b = lerp_idx(xc, Ws, sw);
                    v = a.l0 * (b.l0 * ic[a.i0 * Ws + b.i0] + b.l1 * ic[a.i0 * Ws + b.i1]) +
                        a.l1 * (b.l0 * ic[a.i1 * Ws + b.i0] + b.l1 * ic[a.i1 * Ws + b.i1]);
                } else {
                    v = ic[yy * Ws + xc];
                }
            }
            xs[c][dy][xx] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int xo = 0; xo < Wo; xo++) {
            const float d = dzs[tco][xo];
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int dxx = 0; dxx < 3; dxx++) acc[dy * 3 + dxx] = fmaf(d, xs[tci][dy][xo * S + dxx], acc[dy * 3 + dxx]);
        }
        __syncthreads();
    }
    if (co0 + tco < Cout && ci0 + tci < Cin) {
        float* p = part + (((size_t)blockIdx.z * Cout + co0 + tco) * Cin + ci0 + tci) * 9;
#pragma unroll
        for (int i = 0; i < 9; i++) p[i] = acc[i];
    }
}

// out[i] = scale * sum_k part[k][i]   (fixed order)
__global__ void tr_reduce_kernel(const float* __restrict__ part, int K, long long n, float scale, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < K; k++) s += part[(size_t)k * n + i];
        out[i] = scale * s;
    }
}

// bias gradient: db[c] = scale * sum_{n,pix} dz[n][c][pix]; one CTA per channel, fixed-order tree
__global__ void tr_bias_grad_kernel(const float* __restrict__ dz, int N, int C, int P, float scale, float* __restrict__ out) {
    const int c = blockIdx.x;
    float s = 0.f;
    for (long long i = threadIdx.x; i < (long long)N * P; i += blockDim.x) {
        const int n = (int)(i / P), p = (int)(i - (long long)n * P);
        s += dz[((size_t)n * C + c) * P + p];
    }
    __shared__ float red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[i];
        out[c] = scale * t;
    }
}

// LayerNorm backward (one warp per frame): dz = rstd * (dg*w - mean(dg*w) - xhat * mean(dg*w*xhat)); per-frame dw, db terms
__global__ void tr_ln_bwd_kernel(const float* __restrict__ z, const float* __restrict__ stat, const float* __restrict__ lw,
                                 const float* __restrict__ dg, float* __restrict__ dz, float* __restrict__ dwb, int N, int E) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N) return;
    const float mean = stat[2 * warp], rstd = stat[2 * warp + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < E; c += 32) {
        const float xh = (z[(size_t)warp * E + c] - mean) * rstd, gw = dg[(size_t)warp * E + c] * __ldg(lw + c);
        s1 += gw;
        s2 = fmaf(gw, xh, s2);
    }
    s1 = warp_sum(s1) / (float)E;
    s2 = warp_sum(s2) / (float)E;
    for (int c = lane; c < E; c += 32) {
        const float xh = (z[(size_t)warp * E + c] - mean) * rstd, d = dg[(size_t)warp * E + c];
        dz[(size_t)warp * E + c] = rstd * (d * __ldg(lw + c) - s1 - xh * s2);
        dwb[((size_t)warp * 2) * E + c] = d * xh;
        dwb[((size_t)warp * 2 + 1) * E + c] = d;
    }
}

// Koopman chain + viscosity nets backward, one CTA per sample.  dgp [T-1][B][E]: gradients of the predicted embeddings
// (from the masked decodes).  Outputs: dg0[b][E] (added to the encoder gradient of frame (b, 0)) and this sample's
// gradients of the three nets, laid out like the parameters (w0[50], b0[50], w2[n][50], b2[n]) x 3 in netpart[b][...].
__global__ void tr_koop_bwd_kernel(KoopNets k, const float* __restrict__ visc, const float* __restrict__ G, const float* __restrict__ gpred,
                                   const float* __restrict__ hid, const float* __restrict__ bands, const float* __restrict__ dgp,
                                   float* __restrict__ dg0, float* __restrict__ netpart, int B, int T, int E, float reg2) {
    extern __shared__ float sm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nband = 4 * E - 10, nb = E + 2 * nband;
    float* band = sm;               // [nb]
    float* dband = band + nb;       // [nb]
    float* dcur = dband + nb;       // [E]
    float* gprev = dcur + E;        // [E]
    float* carry = gprev + E;       // [E]
    float* dh = carry + E;          // [150]
    for (int i = tid; i < nb; i += blockDim.x) {
        band[i] = bands[(size_t)b * nb + i];
        dband[i] = reg2 * band[i];                      // d/dK of w_reg * (T-1) * sum K^2
    }
    for (int i = tid; i < E; i += blockDim.x) carry[i] = 0.f;
    __syncthreads();
    const float* D = band;
    const float* U = band + E;
    const float* Lw = U + nband;
    float* dD = dband;
    float* dU = dband + E;
    float* dL = dU + nband;
    int off[NBAND + 1];
    off[1] = 0;
    for (int o = 2; o <= NBAND; o++) off[o] = off[o - 1] + (E - (o - 1));
    for (int t = T - 1; t >= 1; t--) {
        for (int i = tid; i < E; i += blockDim.x) {
            dcur[i] = dgp[((size_t)(t - 1) * B + b) * E + i] + carry[i];
            gprev[i] = (t == 1) ? G[(size_t)(b * T) * E + i] : gpred[((size_t)(t - 2) * B + b) * E + i];
        }
        __syncthreads();
        // g_t[i] = sum_o L_o[i-o] g[i-o] + D[i] g[i] + sum_o U_o[i] g[i+o]
        for (int i = tid; i < E; i += blockDim.x) {
            dD[i] = fmaf(dcur[i], gprev[i], dD[i]);
            float c = D[i] * dcur[i];
            for (int o = 1; o <= NBAND; o++) {
                if (i + o < E) {
                    dU[off[o] + i] = fmaf(dcur[i], gprev[i + o], dU[off[o] + i]);         // row i uses U_o[i] g[i+o]
                    dL[off[o] + i] = fmaf(dcur[i + o], gprev[i], dL[off[o] + i]);         // row i+o uses L_o[i] g[i]
                    c = fmaf(Lw[off[o] + i], dcur[i + o], c);                             // K^T: column i of row i+o
                }
                if (i - o >= 0) c = fmaf(U[off[o] + i - o], dcur[i - o], c);              // column i of row i-o
            }
            carry[i] = c;
        }
        __syncthreads();
    }
    for (int i = tid; i < E; i += blockDim.x) dg0[(size_t)b * E + i] = carry[i];
    // nets: band = W2 h + b2, h = relu(w0 * v + b0)
    const float v = 100.f * visc[b];
    float* np = netpart + (size_t)b * (3 * 100 + nb * 51);
    size_t o2[3];
    o2[0] = 100; o2[1] = 100 + (size_t)E * 51 + 100; o2[2] = o2[1] + (size_t)nband * 51 + 100;
    if (tid < 150) {
        const int net = tid / 50, j = tid % 50;
        const int n = net == 0 ? E : nband;
        const float* db = net == 0 ? dD : (net == 1 ? dU : dL);
        float s = 0.f;
        for (int r = 0; r < n; r++) s = fmaf(db[r], __ldg(k.w2[net] + (size_t)r * 50 + j), s);
        const float hj = hid[(size_t)b * 150 + tid];
        const float dpre = hj > 0.f ? s : 0.f;
        dh[tid] = dpre;
        float* base = np + (net == 0 ? 0 : (net == 1 ? o2[1] - 100 : o2[2] - 100));
        base[j] = dpre * v;           // w0
        base[50 + j] = dpre;          // b0
    }
    __syncthreads();
    for (int i = tid; i < nb; i += blockDim.x) {
        const int net = i < E ? 0 : (i < E + nband ? 1 : 2);
        const int r = i < E ? i : (i < E + nband ? i - E : i - E - nband);
        const int n = net == 0 ? E : nband;
        float* w2g = np + o2[net];
        const float d = dband[i];
        for (int j = 0; j < 50; j++) w2g[(size_t)r * 50 + j] = d * hid[(size_t)b * 150 + net * 50 + j];
        w2g[(size_t)n * 50 + r] = d;   // b2
    }
}

__global__ void tr_add_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows, int E, int dst_stride_rows) {
    // dst[(r * dst_stride_rows) * E + c] += src[r * E + c]
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * E; i += gridDim.x * blockDim.x) {
        const int r = i / E, c = i - r * E;
        dst[((size_t)r * dst_stride_rows) * E + c] += src[i];
    }
}

inline int nblocks(long long n, int per = 256, int cap = 148 * 32) { return (int)std::max<long long>(1, std::min<long long>((n + per - 1) / per, cap)); }

struct Layer {
    int cin, cout, hs, ws, stride;      // hs, ws: layer input size; output = (up ? 2x : 1x) / stride
    bool up, relu;
    int t_w;                            // tensor index of the weight (bias = t_w + 1)
};

}  // namespace

extern "C" {

long long trpx_cyl_num_params(trpx_cyl_t h) {
    if (!cyl_valid(h)) return -1;
    size_t sz[36];
    cyl_tensor_sizes(h->E, sz);
    long long n = 0;
    for (int i = 2; i < 36; i++) n += (long long)sz[i];
    return n;
}

int trpx_cyl_train_fwd_bwd(trpx_cyl_t h, const float* states, const float* visc, int B, int T, float w_rec, float w_dyn,
                           float w_reg, float grad_scale, float* loss2, float* grad, void* stream) {
    TRPX_REQUIRE(cyl_valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "weights not loaded");
    TRPX_REQUIRE(B >= 1 && T >= 1 && states && visc && loss2 && grad, "bad arguments");
    TRPX_REQUIRE((long long)B * (2 * T - 1) <= 16384, "B*(2T-1) = %lld decodes exceed the 16384 of one pass", (long long)B * (2 * T - 1));
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int E = h->E, NF = B * T, ND = B * (2 * T - 1);
    const int nband = 4 * E - 10, nb = E + 2 * nband;
    const float* W = h->blob;
    size_t sz[36], goff[36];
    cyl_tensor_sizes(E, sz);
    {
        size_t o = 0;
        for (int i = 2; i < 36; i++) { goff[i] = o; o += sz[i]; }
    }
    const Layer enc[5] = {{4, 16, 64, 128, 2, false, true, T_ENC + 0}, {16, 32, 32, 64, 2, false, true, T_ENC + 2},
                          {32, 64, 16, 32, 2, false, true, T_ENC + 4}, {64, 128, 8, 16, 2, false, true, T_ENC + 6},
                          {128, E / 32, 4, 8, 1, false, false, T_ENC + 8}};
    const Layer dec[5] = {{E / 32, 128, 4, 8, 1, true, true, T_DEC + 0}, {128, 64, 8, 16, 1, true, true, T_DEC + 2},
                          {64, 32, 16, 32, 1, true, true, T_DEC + 4}, {32, 16, 32, 64, 1, true, true, T_DEC + 6},
                          {16, 3, 64, 128, 1, false, false, T_DEC + 8}};
    auto out_elems = [](const Layer& l) {
        const int hi = l.up ? 2 * l.hs : l.hs, wi = l.up ? 2 * l.ws : l.ws;
        return (size_t)l.cout * ((hi - 1) / l.stride + 1) * ((wi - 1) / l.stride + 1);
    };
    // ---- workspace carve-up (floats) ----
    size_t o = 0;
    auto take = [&](size_t n) { const size_t r = o; o += (n + 63) & ~(size_t)63; return r; };
    size_t ea[6], eg[5], da[6], dgd[6];
    ea[0] = take((size_t)NF * 4 * HW);
    for (int l = 0; l < 5; l++) ea[l + 1] = take((size_t)NF * out_elems(enc[l]));
    const size_t o_stat = take((size_t)NF * 2), o_G = take((size_t)ND * E), o_hid = take((size_t)B * 150),
                 o_bands = take((size_t)B * nb), o_ksq = take(B);
    da[0] = o_G;
    for (int l = 0; l < 5; l++) da[l + 1] = take((size_t)ND * out_elems(dec[l]));
    for (int l = 0; l < 5; l++) dgd[l + 1] = take((size_t)ND * out_elems(dec[l]));       // gradient at each decoder layer output
    dgd[0] = take((size_t)ND * E);                                                       // gradient at the decoder input
    const size_t o_dU = take((size_t)ND * 32 * HW);                                      // largest conv-input gradient (upsampled res)
    for (int l = 0; l < 5; l++) eg[l] = take((size_t)NF * out_elems(enc[l]));            // gradient at each encoder layer output
    const size_t o_dgenc = take((size_t)NF * E), o_dg0 = take((size_t)B * E), o_lnwb = take((size_t)NF * 2 * E);
    const int loss_per = 24;
    const size_t o_lpart = take((size_t)ND * loss_per);
    const size_t netsz = 3 * 100 + (size_t)nb * 51;
    const size_t o_net = take((size_t)B * netsz);
    const int ksplit = 96;
    const size_t o_part = take((size_t)ksplit * 128 * 64 * 9);
    if (o * sizeof(float) > h->tws_bytes) {
        if (h->tws) { TRPX_CUDA(cudaDeviceSynchronize()); TRPX_CUDA(cudaFree(h->tws)); h->tws = nullptr; h->tws_bytes = 0; }
        TRPX_CUDA(cudaMalloc(&h->tws, o * sizeof(float)));
        h->tws_bytes = o * sizeof(float);
    }
    float* ws = h->tws;
    auto conv_fwd = [&](const Layer& l, const float* in, float* out, int N) {
        const long long total = (long long)N * out_elems(l);
        if (l.up) tr_conv_fwd_kernel<true><<<nblocks(total), 256, 0, st>>>(in, W + h->off[l.t_w], W + h->off[l.t_w + 1], out, N, l.cin, l.hs, l.ws, l.cout, l.stride, l.relu);
        else tr_conv_fwd_kernel<false><<<nblocks(total), 256, 0, st>>>(in, W + h->off[l.t_w], W + h->off[l.t_w + 1], out, N, l.cin, l.hs, l.ws, l.cout, l.stride, l.relu);
    };
    // ================= forward =================
    tr_prep_kernel<<<nblocks((long long)NF * 4 * HW), 256, 0, st>>>(states, visc, W + h->off[T_MU], W + h->off[T_STD], ws + ea[0], NF, T);
    for (int l = 0; l < 5; l++) conv_fwd(enc[l], ws + ea[l], ws + ea[l + 1], NF);
    tr_ln_fwd_kernel<<<(NF * 32 + 255) / 256, 256, 0, st>>>(ws + ea[5], W + h->off[T_LNW], W + h->off[T_LNB], ws + o_G, ws + o_stat, NF, E, h->eps);
    KoopNets kn;
    const int tk[3] = {T_KD, T_KU, T_KL};
    for (int i = 0; i < 3; i++) {
        kn.w0[i] = W + h->off[tk[i]]; kn.b0[i] = W + h->off[tk[i] + 1]; kn.w2[i] = W + h->off[tk[i] + 2]; kn.b2[i] = W + h->off[tk[i] + 3];
    }
    {
        const size_t smem = sizeof(float) * (150 + nb + 2 * E + 32);
        tr_koop_fwd_kernel<<<B, 256, smem, st>>>(kn, visc, ws + o_G, ws + o_hid, ws + o_bands, ws + o_ksq, ws + o_G + (size_t)NF * E, B, T, E);
    }
    for (int l = 0; l < 5; l++) conv_fwd(dec[l], ws + da[l], ws + da[l + 1], ND);
    TRPX_CUDA(cudaGetLastError());
    // ================= loss =================
    const float norm = 1.0f / ((float)B * 3.0f * (float)HW);
    tr_loss_grad_kernel<<<dim3(loss_per, ND), 256, 0, st>>>(ws + da[5], states, W + h->off[T_MU], W + h->off[T_STD], h->mask, ws + dgd[5],
                                                           ws + o_lpart, NF, B, T, w_rec * norm, w_dyn * norm);
    tr_loss_final_kernel<<<1, 32, 0, st>>>(ws + o_lpart, loss_per, NF, ND, ws + o_ksq, B, T, w_rec, w_dyn, w_reg, loss2);
    // ================= backward =================
    auto conv_bwd = [&](const Layer& l, const float* in, float* gout /* gradient at the layer output, ReLU already applied */,
                        float* gin /* gradient at the layer input, or null */, int N) -> int {
        const int hi = l.up ? 2 * l.hs : l.hs, wi = l.up ? 2 * l.ws : l.ws;
        const int ho = (hi - 1) / l.stride + 1, wo = (wi - 1) / l.stride + 1;
        TRPX_REQUIRE(wo <= BW_MAXW, "row too wide");
        dim3 grid((l.cout + BW_CO - 1) / BW_CO, (l.cin + BW_CI - 1) / BW_CI, std::min(ksplit, N * ho));
        float* part = ws + o_part;
        if (l.up) tr_conv_bwd_weight_kernel<1, true><<<grid, 256, 0, st>>>(in, gout, part, N, l.cin, l.hs, l.ws, l.cout, ho, wo);
        else if (l.stride == 2) tr_conv_bwd_weight_kernel<2, false><<<grid, 256, 0, st>>>(in, gout, part, N, l.cin, l.hs, l.ws, l.cout, ho, wo);
        else tr_conv_bwd_weight_kernel<1, false><<<grid, 256, 0, st>>>(in, gout, part, N, l.cin, l.hs, l.ws, l.cout, ho, wo);
        const long long nw = (long long)l.cout * l.cin * 9;
        tr_reduce_kernel<<<nblocks(nw), 256, 0, st>>>(part, (int)grid.z, nw, grad_scale, grad + goff[l.t_w]);
        tr_bias_grad_kernel<<<l.cout, 256, 0, st>>>(gout, N, l.cout, ho * wo, grad_scale, grad + goff[l.t_w + 1]);
        if (gin) {
            float* dxi = l.up ? ws + o_dU : gin;       // gradient at the conv-input resolution
            const long long total = (long long)N * l.cin * hi * wi;
            if (l.stride == 2) tr_conv_bwd_data_kernel<2><<<nblocks(total), 256, 0, st>>>(gout, W + h->off[l.t_w], dxi, N, l.cin, hi, wi, l.cout, ho, wo);
            else tr_conv_bwd_data_kernel<1><<<nblocks(total), 256, 0, st>>>(gout, W + h->off[l.t_w], dxi, N, l.cin, hi, wi, l.cout, ho, wo);
            if (l.up) tr_upsample_bwd_kernel<<<nblocks((long long)N * l.cin * l.hs * l.ws), 256, 0, st>>>(dxi, gin, (long long)N * l.cin, l.hs, l.ws);
        }
        TRPX_CUDA(cudaGetLastError());
        return TRPX_OK;
    };
    int rc;
    for (int l = 4; l >= 0; l--) {
        if (dec[l].relu) {
            const long long n = (long long)ND * out_elems(dec[l]);
            tr_relu_bwd_kernel<<<nblocks(n), 256, 0, st>>>(ws + dgd[l + 1], ws + da[l + 1], n);
        }
        if ((rc = conv_bwd(dec[l], ws + da[l], ws + dgd[l + 1], ws + dgd[l], ND))) return rc;
    }
    // Koopman chain and nets (gradients of the predicted embeddings are rows NF.. of the decoder-input gradient)
    {
        const size_t smem = sizeof(float) * (2 * nb + 3 * E + 150);
        tr_koop_bwd_kernel<<<B, 256, smem, st>>>(kn, visc, ws + o_G, ws + o_G + (size_t)NF * E, ws + o_hid, ws + o_bands,
                                                 ws + dgd[0] + (size_t)NF * E, ws + o_dg0, ws + o_net, B, T, E,
                                                 2.0f * w_reg * (float)(T - 1));
        // per-sample partials -> flat gradient: the three nets are contiguous in the parameter order
        tr_reduce_kernel<<<nblocks((long long)netsz), 256, 0, st>>>(ws + o_net, B, (long long)netsz, grad_scale, grad + goff[T_KD]);
    }
    // encoder: dG = decoder-input gradient of the reconstruction decodes (+ chain gradient for the t = 0 frames)
    TRPX_CUDA(cudaMemcpyAsync(ws + o_dgenc, ws + dgd[0], (size_t)NF * E * sizeof(float), cudaMemcpyDeviceToDevice, st));
    tr_add_kernel<<<nblocks((long long)B * E), 256, 0, st>>>(ws + o_dgenc, ws + o_dg0, B, E, T);
    tr_ln_bwd_kernel<<<(NF * 32 + 255) / 256, 256, 0, st>>>(ws + ea[5], ws + o_stat, W + h->off[T_LNW], ws + o_dgenc, ws + eg[4], ws + o_lnwb, NF, E);
    tr_reduce_kernel<<<nblocks(2 * E), 256, 0, st>>>(ws + o_lnwb, NF, 2 * E, grad_scale, grad + goff[T_LNW]);
    for (int l = 4; l >= 0; l--) {
        if (enc[l].relu) {
            const long long n = (long long)NF * out_elems(enc[l]);
            tr_relu_bwd_kernel<<<nblocks(n), 256, 0, st>>>(ws + eg[l], ws + ea[l + 1], n);
        }
        if ((rc = conv_bwd(enc[l], ws + ea[l], ws + eg[l], l > 0 ? ws + eg[l - 1] : nullptr, NF))) return rc;
    }
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

}  // extern "C"
