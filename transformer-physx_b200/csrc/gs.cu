// Gray-Scott Koopman embedding: trphysx/embedding/embedding_grayscott.py:25-218 in eval mode.
//
//   embed   : (x - mu) / std -> [Conv3d circular (5^3 s2, 3^3 s2, 3^3 s2, 1^3) + BatchNorm3d (running stats) + LeakyReLU] x 4
//             -> Linear(4096, 512) + LeakyReLU -> Linear(512, E) -> LayerNorm(E)                              (:46-66,127-139)
//   recover : Linear(E, 512) -> Linear(512, 4096) + LeakyReLU -> Conv 1^3 + BN + LReLU -> 3 x [trilinear x2
//             (align_corners=False) -> Conv3d circular 3^3 + BN + LReLU] -> Conv 1^3 (64 -> 2) -> std * y + mu   (:68-98,141-153)
//   koopman : skew-symmetric banded operator (9 off-diagonals) + diagonal, as a banded matvec; dense on demand   (:155-175)
//
// FP32 SIMT throughout (the row is a compatibility row, not a headline path).  The convs are direct convolutions with the
// BatchNorm affine folded into a per-channel scale/shift at load time: a CTA owns a 4x4x4 block of output voxels x 32
// output channels, stages an input patch (circular wrap resolved while staging) and the weight slice of 8 input channels
// in shared memory, and each thread accumulates 8 output channels of one voxel (one patch word + two broadcast LDS.128
// per 8 FMAs).
#include "common.cuh"
#include <algorithm>
#include <vector>

struct trpx_gs {
    uint32_t magic = 0x67737774u;
    int E = 512;
    float ln_eps = 1e-5f;
    int device = 0;
    bool loaded = false;
    float* blob = nullptr;        // packed weights (transposed conv kernels, folded BN, FC matrices)
    size_t blob_floats = 0;
    float* ws = nullptr;          // activations of the chunk in flight
    size_t ws_bytes = 0;
};

namespace trpx {
namespace {

inline bool valid(trpx_gs_t h) { return h != nullptr && h->magic == 0x67737774u; }

struct ConvSpec { int cin, cout, k, s, din; float slope; };      // din: input edge length; output edge = din / s
// encoder convs 0..3, decoder convs 4..7 (recoveryNet.0 / .4 / .8 / .12); recoveryNet.15 (64 -> 2) is handled apart
constexpr ConvSpec kConv[8] = {{2, 64, 5, 2, 32, 0.02f},  {64, 128, 3, 2, 16, 0.02f}, {128, 128, 3, 2, 8, 0.02f},
                               {128, 64, 1, 1, 4, 0.02f}, {64, 128, 1, 1, 4, 0.02f},  {128, 128, 3, 1, 8, 0.02f},
                               {128, 64, 3, 1, 16, 0.02f}, {64, 64, 3, 1, 32, 0.02f}};

struct GsLayout {
    // per conv: transposed weights [cin][k^3][cout], scale[cout], shift[cout]
    size_t wT[8], scale[8], shift[8];
    size_t fc1w, fc1b, fc2w, fc2b, lnw, lnb;            // observableNetFC.0 / .2 / .3
    size_t rf1w, rf1b, rf2w, rf2b;                      // recoveryNetFC.0 / .2
    size_t w15, b15;                                    // recoveryNet.15  [2][64], [2]
    size_t kdiag, kut, mu, std, total;
    static GsLayout make(int E) {
        GsLayout L{};
        size_t o = 0;
        auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~size_t(3); return r; };
        for (int i = 0; i < 8; i++) {
            const ConvSpec& c = kConv[i];
            L.wT[i] = take((size_t)c.cin * c.k * c.k * c.k * c.cout);
            L.scale[i] = take(c.cout);
            L.shift[i] = take(c.cout);
        }
        L.fc1w = take((size_t)512 * 4096); L.fc1b = take(512);
        L.fc2w = take((size_t)E * 512); L.fc2b = take(E);
        L.lnw = take(E); L.lnb = take(E);
        L.rf1w = take((size_t)512 * E); L.rf1b = take(512);
        L.rf2w = take((size_t)4096 * 512); L.rf2b = take(4096);
        L.w15 = take(128); L.b15 = take(2);
        L.kdiag = take(E);
        size_t nb = 0;
        for (int d = 1; d < 10; d++) nb += E - d;
        L.kut = take(nb);
        L.mu = take(1); L.std = take(1);
        L.total = o;
        return L;
    }
};

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }

// ---- weight preparation ----
__global__ void gs_prep_conv_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ bnw,
                                    const float* __restrict__ bnb, const float* __restrict__ mean, const float* __restrict__ var,
                                    float* __restrict__ wT, float* __restrict__ scale, float* __restrict__ shift, int cin, int cout, int k3) {
    const int total = cin * k3 * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % cout, r = i / cout, tap = r % k3, ci = r / k3;
        wT[i] = w[((size_t)co * cin + ci) * k3 + tap];
    }
    for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < cout; co += gridDim.x * blockDim.x) {
        const float sc = bnw[co] / sqrtf(var[co] + 1e-5f);          // nn.BatchNorm3d default eps
        scale[co] = sc;
        shift[co] = (b[co] - mean[co]) * sc + bnb[co];
    }
}

// ---- direct 3-D convolution, circular padding, fused scale/shift + LeakyReLU ----
template <int K, int S>
__global__ void __launch_bounds__(256) gs_conv3d_kernel(const float* __restrict__ in, const float* __restrict__ wT,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        float* __restrict__ out, int cin, int cout, int din, float slope,
                                                        const float* __restrict__ mu, const float* __restrict__ sd) {
    constexpr int P = 3 * S + K, P3 = P * P * P, K3 = K * K * K, PAD = K / 2;
    constexpr int CC = (K == 5) ? 2 : 8;
    extern __shared__ __align__(16) float smem[];
    float* patch = smem;                       // [CC][P3]
    float* wsm = smem + ((CC * P3 + 3) & ~3);  // [CC][K3][32], 16-byte aligned
    const int dout = din / S, tpd = dout / 4;
    const int tile = blockIdx.x, tz = tile / (tpd * tpd), ty = (tile / tpd) % tpd, tx = tile % tpd;
    const int co0 = blockIdx.y * 32, n = blockIdx.z;
    const int tid = threadIdx.x, vox = tid & 63, cg = tid >> 6;
    const int vz = vox >> 4, vy = (vox >> 2) & 3, vx = vox & 3;
    const float nm = mu ? mu[0] : 0.f;
    const bool norm = (mu != nullptr);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
    const float* inb = in + (size_t)n * cin * din * din * din;
    const int z0 = tz * 4 * S - PAD, y0 = ty * 4 * S - PAD, x0 = tx * 4 * S - PAD;
    for (int c0 = 0; c0 < cin; c0 += CC) {
        __syncthreads();
        for (int i = tid; i < CC * P3; i += 256) {
            const int ci = i / P3, r = i % P3, pz = r / (P * P), py = (r / P) % P, px = r % P;
            const int z = (z0 + pz + din) % din, y = (y0 + py + din) % din, x = (x0 + px + din) % din;
            float v = inb[(((size_t)(c0 + ci) * din + z) * din + y) * din + x];
            if (norm) v = (v - nm) / sd[0];
            patch[i] = v;
        }
        for (int i = tid; i < CC * K3 * 32; i += 256) {
            const int co = i & 31, r = i >> 5;          // r = ci * K3 + tap
            wsm[i] = wT[((size_t)c0 * K3 + r) * cout + co0 + co];
        }
        __syncthreads();
#pragma unroll 1
        for (int ci = 0; ci < CC; ci++) {
            const float* pp = patch + ci * P3 + (vz * S) * P * P + (vy * S) * P + vx * S;
            const float* ww = wsm + (ci * K3) * 32 + cg * 8;
#pragma unroll
            for (int kz = 0; kz < K; kz++)
#pragma unroll
                for (int ky = 0; ky < K; ky++)
#pragma unroll
                    for (int kx = 0; kx < K; kx++) {
                        const float a = pp[kz * P * P + ky * P + kx];
                        const float4 w0 = *reinterpret_cast<const float4*>(ww + ((kz * K + ky) * K + kx) * 32);
                        const float4 w1 = *reinterpret_cast<const float4*>(ww + ((kz * K + ky) * K + kx) * 32 + 4);
                        acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
                        acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
                        acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
                        acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
                    }
        }
    }
    const int oz = tz * 4 + vz, oy = ty * 4 + vy, ox = tx * 4 + vx;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int co = co0 + cg * 8 + i;
        const float y = lrelu(fmaf(acc[i], scale[co], shift[co]), slope);
        out[((((size_t)n * cout + co) * dout + oz) * dout + oy) * dout + ox] = y;
    }
}

// ---- trilinear x2, align_corners=False (aten upsample_trilinear3d: source index clamped at 0, edge clamp above) ----
__global__ void gs_upsample2_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int d) {
    const int D = 2 * d;
    const size_t total = (size_t)planes * D * D * D;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ox = (int)(i % D), oy = (int)((i / D) % D), oz = (int)((i / ((size_t)D * D)) % D);
        const size_t pl = i / ((size_t)D * D * D);
        int i0[3], i1[3];
        float l0[3], l1[3];
        const int o3[3] = {oz, oy, ox};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float src = 0.5f * ((float)o3[a] + 0.5f) - 0.5f;
            if (src < 0.f) src = 0.f;
            i0[a] = (int)src;
            i1[a] = i0[a] + ((i0[a] < d - 1) ? 1 : 0);
            l1[a] = src - (float)i0[a];
            l0[a] = 1.0f - l1[a];
        }
        const float* p = in + pl * d * d * d;
        auto at = [&](int z, int y, int x) { return p[((size_t)z * d + y) * d + x]; };
        const float v = l0[0] * (l0[1] * (l0[2] * at(i0[0], i0[1], i0[2]) + l1[2] * at(i0[0], i0[1], i1[2])) +
                                 l1[1] * (l0[2] * at(i0[0], i1[1], i0[2]) + l1[2] * at(i0[0], i1[1], i1[2]))) +
                        l1[0] * (l0[1] * (l0[2] * at(i1[0], i0[1], i0[2]) + l1[2] * at(i1[0], i0[1], i1[2])) +
                                 l1[1] * (l0[2] * at(i1[0], i1[1], i0[2]) + l1[2] * at(i1[0], i1[1], i1[2])));
        out[i] = v;
    }
}

// ---- Linear: out[n][m] = act(in[n][:] . W[m][:] + b[m]); one warp per output m, 8 samples per pass ----
__global__ void __launch_bounds__(256) gs_linear_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                        const float* __restrict__ b, float* __restrict__ out, int N, int K, int M,
                                                        float slope) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * 8 + warp;
    const int n0 = blockIdx.y * 8;
    if (m >= M) return;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.f;
    const float* wr = W + (size_t)m * K;
    for (int k = lane * 4; k < K; k += 128) {
        const float4 w4 = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (n0 + i < N) {
                const float4 a = *reinterpret_cast<const float4*>(in + (size_t)(n0 + i) * K + k);
                acc[i] = fmaf(a.x, w4.x, acc[i]); acc[i] = fmaf(a.y, w4.y, acc[i]);
                acc[i] = fmaf(a.z, w4.z, acc[i]); acc[i] = fmaf(a.w, w4.w, acc[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = warp_sum(acc[i]);
    if (lane < 8 && n0 + lane < N) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) if (lane == i) v = acc[i];
        out[(size_t)(n0 + lane) * M + m] = lrelu(v + b[m], slope);
    }
}

__global__ void gs_layernorm_kernel(float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int N, int E,
                                    float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    float* r = x + (size_t)row * E;
    float s = 0.f;
    for (int i = lane; i < E; i += 32) s += r[i];
    const float mean = warp_sum(s) / (float)E;
    float v = 0.f;
    for (int i = lane; i < E; i += 32) { const float d = r[i] - mean; v = fmaf(d, d, v); }
    const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)E + eps);
    for (int i = lane; i < E; i += 32) r[i] = (r[i] - mean) * rstd * w[i] + b[i];
}

// ---- last layer: Conv3d 1^3 (64 -> 2) + un-normalise; thread per voxel ----
__global__ void gs_final_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ b,
                                const float* __restrict__ mu, const float* __restrict__ sd, float* __restrict__ out, int N) {
    const int V = 32 * 32 * 32;
    const size_t total = (size_t)N * V;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t n = i / V, v = i % V;
        const float* p = in + n * 64 * V + v;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
        for (int c = 0; c < 64; c++) {
            const float x = p[(size_t)c * V];
            a0 = fmaf(x, __ldg(w + c), a0);
            a1 = fmaf(x, __ldg(w + 64 + c), a1);
        }
        out[(n * 2 + 0) * V + v] = sd[0] * (a0 + b[0]) + mu[0];
        out[(n * 2 + 1) * V + v] = sd[0] * (a1 + b[1]) + mu[0];
    }
}

// ---- Koopman operator: K[i][i+d] = ut_d[i], K[i+d][i] = -ut_d[i] (d = 1..9), K[i][i] = diag[i] ----
__global__ void gs_koopman_kernel(const float* __restrict__ g, const float* __restrict__ diag, const float* __restrict__ ut,
                                  float* __restrict__ out, float* __restrict__ kmat, int B, int E) {
    const size_t total = (size_t)B * E;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % E);
        const float* gr = g + (idx - i);
        float* kr = kmat ? kmat + idx * E : nullptr;
        if (kr) {
            for (int j = 0; j < E; j++) kr[j] = 0.f;
            kr[i] = diag[i];
        }
        // the dense bmm of the reference sums the row left to right; keep that order
        float acc = 0.f;
        int off[10];
        off[1] = 0;
#pragma unroll
        for (int d = 2; d < 10; d++) off[d] = off[d - 1] + (E - (d - 1));
#pragma unroll
        for (int d = 9; d >= 1; d--)
            if (i - d >= 0) {
                const float kv = -ut[off[d] + i - d];
                acc = fmaf(kv, gr[i - d], acc);
                if (kr) kr[i - d] = kv;
            }
        acc = fmaf(diag[i], gr[i], acc);
#pragma unroll
        for (int d = 1; d < 10; d++)
            if (i + d < E) {
                const float kv = ut[off[d] + i];
                acc = fmaf(kv, gr[i + d], acc);
                if (kr) kr[i + d] = kv;
            }
        out[idx] = acc;
    }
}

template <int K, int S>
int launch_conv(const ConvSpec& c, const float* in, const float* blob, const GsLayout& L, int li, float* out, int N,
                bool normalise, cudaStream_t st) {
    constexpr int P = 3 * S + K, CC = (K == 5) ? 2 : 8;
    const size_t smem = ((size_t)((CC * P * P * P + 3) & ~3) + (size_t)CC * K * K * K * 32) * sizeof(float);
    TRPX_CUDA(cudaFuncSetAttribute(gs_conv3d_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int dout = c.din / S, tpd = dout / 4;
    dim3 grid(tpd * tpd * tpd, c.cout / 32, N);
    gs_conv3d_kernel<K, S><<<grid, 256, smem, st>>>(in, blob + L.wT[li], blob + L.scale[li], blob + L.shift[li], out, c.cin,
                                                     c.cout, c.din, c.slope, normalise ? blob + L.mu : nullptr,
                                                     normalise ? blob + L.std : nullptr);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

int run_conv(int li, const float* in, const float* blob, const GsLayout& L, float* out, int N, cudaStream_t st) {
    const ConvSpec& c = kConv[li];
    if (c.k == 5) return launch_conv<5, 2>(c, in, blob, L, li, out, N, true, st);
    if (c.k == 3 && c.s == 2) return launch_conv<3, 2>(c, in, blob, L, li, out, N, false, st);
    if (c.k == 3) return launch_conv<3, 1>(c, in, blob, L, li, out, N, false, st);
    return launch_conv<1, 1>(c, in, blob, L, li, out, N, false, st);
}

int ensure_ws(trpx_gs_t h, size_t bytes) {
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

int linear(const float* in, const float* W, const float* b, float* out, int N, int K, int M, float slope, cudaStream_t st) {
    dim3 grid((M + 7) / 8, (N + 7) / 8);
    gs_linear_kernel<<<grid, 256, 0, st>>>(in, W, b, out, N, K, M, slope);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

constexpr int GS_CHUNK = 16;                                  // frames per pass: bounds the workspace (64 ch x 32^3 x 2 buffers)
constexpr size_t GS_BUF = (size_t)64 * 32 * 32 * 32;          // floats per frame of the largest intermediate (64 channels x 32^3)

// decoder from the [N, 64, 4, 4, 4] latent volume
int decode_latent(trpx_gs_t h, const float* z, float* x, int N, cudaStream_t st) {
    const GsLayout L = GsLayout::make(h->E);
    for (int n0 = 0; n0 < N; n0 += GS_CHUNK) {
        const int nn = std::min(GS_CHUNK, N - n0);
        int rc = ensure_ws(h, 2 * GS_BUF * GS_CHUNK * sizeof(float));
        if (rc) return rc;
        float* a = h->ws;
        float* b = h->ws + GS_BUF * GS_CHUNK;
        const float* zin = z + (size_t)n0 * 4096;
        if ((rc = run_conv(4, zin, h->blob, L, a, nn, st))) return rc;                       // 64 -> 128 @ 4^3
        int d = 4;
        const int chans[3] = {128, 128, 64};
        for (int s = 0; s < 3; s++) {
            const size_t tot = (size_t)nn * chans[s] * 8 * d * d * d;
            gs_upsample2_kernel<<<(unsigned)std::min<size_t>((tot + 255) / 256, 65535 * 8), 256, 0, st>>>(a, b, nn * chans[s], d);
            TRPX_CUDA(cudaGetLastError());
            d *= 2;
            if ((rc = run_conv(5 + s, b, h->blob, L, a, nn, st))) return rc;
        }
        const size_t tot = (size_t)nn * 32768;
        gs_final_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a, h->blob + L.w15, h->blob + L.b15, h->blob + L.mu,
                                                                        h->blob + L.std, x + (size_t)n0 * 2 * 32768, nn);
        TRPX_CUDA(cudaGetLastError());
    }
    return TRPX_OK;
}

}  // namespace
}  // namespace trpx

using namespace trpx;

extern "C" {

int trpx_gs_create(int n_embd, float ln_eps, int device, trpx_gs_t* out) {
    TRPX_REQUIRE(out != nullptr, "null output");
    TRPX_REQUIRE(n_embd >= 32 && n_embd % 4 == 0 && n_embd <= 4096, "n_embd=%d unsupported", n_embd);
    int ndev = 0;
    TRPX_CUDA(cudaGetDeviceCount(&ndev));
    TRPX_REQUIRE(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
    DeviceGuard guard(device);
    trpx_gs* h = new (std::nothrow) trpx_gs();
    TRPX_REQUIRE(h != nullptr, "out of host memory");
    h->E = n_embd;
    h->ln_eps = ln_eps;
    h->device = device;
    const GsLayout L = GsLayout::make(n_embd);
    h->blob_floats = L.total;
    if (cudaMalloc(&h->blob, L.total * sizeof(float)) != cudaSuccess) {
        delete h;
        return set_error(TRPX_E_CUDA, "cudaMalloc of the Gray-Scott weight blob failed");
    }
    *out = h;
    return TRPX_OK;
}

int trpx_gs_destroy(trpx_gs_t h) {
    if (!valid(h)) return set_error(TRPX_E_STATE, "trpx_gs_destroy: invalid handle");
    DeviceGuard guard(h->device);
    cudaFree(h->blob);
    cudaFree(h->ws);
    h->magic = 0;
    delete h;
    return TRPX_OK;
}

/* 2 (mu, std) + 8 convs x 6 (weight, bias, bn.weight, bn.bias, bn.running_mean, bn.running_var) + 6 (observableNetFC)
 * + 4 (recoveryNetFC) + 2 (recoveryNet.15) + 2 (kMatrixDiag, kMatrixUT) */
int trpx_gs_num_weight_tensors(trpx_gs_t h) { return valid(h) ? 2 + 48 + 6 + 4 + 2 + 2 : -1; }

int trpx_gs_load_weights(trpx_gs_t h, const float* const* w, int n, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    TRPX_REQUIRE(w != nullptr && n == 64, "expected 64 weight tensors, got %d", n);
    for (int i = 0; i < n; i++) TRPX_REQUIRE(w[i] != nullptr, "weight tensor %d is null", i);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const GsLayout L = GsLayout::make(h->E);
    auto cp = [&](size_t off, const float* src, size_t cnt) {
        return cudaMemcpyAsync(h->blob + off, src, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st);
    };
    TRPX_CUDA(cp(L.mu, w[0], 1));
    TRPX_CUDA(cp(L.std, w[1], 1));
    for (int i = 0; i < 8; i++) {
        const ConvSpec& c = kConv[i];
        const float* const* t = w + 2 + 6 * i;
        gs_prep_conv_kernel<<<128, 256, 0, st>>>(t[0], t[1], t[2], t[3], t[4], t[5], h->blob + L.wT[i], h->blob + L.scale[i],
                                                 h->blob + L.shift[i], c.cin, c.cout, c.k * c.k * c.k);
        TRPX_CUDA(cudaGetLastError());
    }
    const float* const* f = w + 50;
    const int E = h->E;
    TRPX_CUDA(cp(L.fc1w, f[0], (size_t)512 * 4096)); TRPX_CUDA(cp(L.fc1b, f[1], 512));
    TRPX_CUDA(cp(L.fc2w, f[2], (size_t)E * 512));    TRPX_CUDA(cp(L.fc2b, f[3], E));
    TRPX_CUDA(cp(L.lnw, f[4], E));                   TRPX_CUDA(cp(L.lnb, f[5], E));
    TRPX_CUDA(cp(L.rf1w, f[6], (size_t)512 * E));    TRPX_CUDA(cp(L.rf1b, f[7], 512));
    TRPX_CUDA(cp(L.rf2w, f[8], (size_t)4096 * 512)); TRPX_CUDA(cp(L.rf2b, f[9], 4096));
    TRPX_CUDA(cp(L.w15, f[10], 128));                TRPX_CUDA(cp(L.b15, f[11], 2));
    TRPX_CUDA(cp(L.kdiag, f[12], E));
    size_t nb = 0;
    for (int d = 1; d < 10; d++) nb += E - d;
    TRPX_CUDA(cp(L.kut, f[13], nb));
    h->loaded = true;
    return TRPX_OK;
}

int trpx_gs_embed(trpx_gs_t h, const float* x, float* g, float* g0, int N, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "trpx_gs_embed: weights not loaded");
    TRPX_REQUIRE(N >= 0, "negative size");
    if (N == 0) return TRPX_OK;
    TRPX_REQUIRE(x != nullptr && g != nullptr, "null tensor");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const GsLayout L = GsLayout::make(h->E);
    for (int n0 = 0; n0 < N; n0 += GS_CHUNK) {
        const int nn = std::min(GS_CHUNK, N - n0);
        int rc = ensure_ws(h, 2 * GS_BUF * GS_CHUNK * sizeof(float));
        if (rc) return rc;
        float* a = h->ws;
        float* b = h->ws + GS_BUF * GS_CHUNK;
        if ((rc = run_conv(0, x + (size_t)n0 * 2 * 32768, h->blob, L, a, nn, st))) return rc;      // 2 -> 64, 32^3 -> 16^3
        if ((rc = run_conv(1, a, h->blob, L, b, nn, st))) return rc;                               // 64 -> 128, -> 8^3
        if ((rc = run_conv(2, b, h->blob, L, a, nn, st))) return rc;                               // 128 -> 128, -> 4^3
        float* z = g0 ? g0 + (size_t)n0 * 4096 : b;
        if ((rc = run_conv(3, a, h->blob, L, z, nn, st))) return rc;                               // 128 -> 64 @ 4^3
        if ((rc = linear(z, h->blob + L.fc1w, h->blob + L.fc1b, a, nn, 4096, 512, 0.02f, st))) return rc;
        float* go = g + (size_t)n0 * h->E;
        if ((rc = linear(a, h->blob + L.fc2w, h->blob + L.fc2b, go, nn, 512, h->E, 1.0f, st))) return rc;
        gs_layernorm_kernel<<<(nn + 7) / 8, 256, 0, st>>>(go, h->blob + L.lnw, h->blob + L.lnb, nn, h->E, h->ln_eps);
        TRPX_CUDA(cudaGetLastError());
    }
    return TRPX_OK;
}

int trpx_gs_recover(trpx_gs_t h, const float* g, float* x, float* out0, int N, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "trpx_gs_recover: weights not loaded");
    TRPX_REQUIRE(N >= 0, "negative size");
    if (N == 0) return TRPX_OK;
    TRPX_REQUIRE(g != nullptr && x != nullptr, "null tensor");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const GsLayout L = GsLayout::make(h->E);
    float* z = out0;
    float* ztmp = nullptr;
    if (z == nullptr) {
        TRPX_CUDA(cudaMallocAsync(&ztmp, (size_t)N * (4096 + 512) * sizeof(float), st));
        z = ztmp;
    }
    float* hid = nullptr;
    if (ztmp) hid = ztmp + (size_t)N * 4096;
    else TRPX_CUDA(cudaMallocAsync(&hid, (size_t)N * 512 * sizeof(float), st));
    int rc = linear(g, h->blob + L.rf1w, h->blob + L.rf1b, hid, N, h->E, 512, 1.0f, st);     // LeakyReLU(1.0) = identity
    if (!rc) rc = linear(hid, h->blob + L.rf2w, h->blob + L.rf2b, z, N, 512, 4096, 0.02f, st);
    if (!rc) rc = decode_latent(h, z, x, N, st);
    if (ztmp) cudaFreeAsync(ztmp, st);
    else cudaFreeAsync(hid, st);
    return rc;
}

int trpx_gs_decode_latent(trpx_gs_t h, const float* z, float* x, int N, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "trpx_gs_decode_latent: weights not loaded");
    TRPX_REQUIRE(N >= 0, "negative size");
    if (N == 0) return TRPX_OK;
    TRPX_REQUIRE(z != nullptr && x != nullptr, "null tensor");
    DeviceGuard guard(h->device);
    return decode_latent(h, z, x, N, (cudaStream_t)stream);
}

int trpx_gs_koopman(trpx_gs_t h, const float* g, float* out, float* kmatrix, int B, void* stream) {
    TRPX_REQUIRE(valid(h), "invalid handle");
    if (!h->loaded) return set_error(TRPX_E_STATE, "trpx_gs_koopman: weights not loaded");
    TRPX_REQUIRE(B >= 0, "negative size");
    if (B == 0) return TRPX_OK;
    TRPX_REQUIRE(g != nullptr && out != nullptr, "null tensor");
    DeviceGuard guard(h->device);
    const GsLayout L = GsLayout::make(h->E);
    const size_t total = (size_t)B * h->E;
    gs_koopman_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g, h->blob + L.kdiag, h->blob + L.kut, out,
                                                                                         kmatrix, B, h->E);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

}  // extern "C"
