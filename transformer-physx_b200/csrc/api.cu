// libtrpx: error plumbing, version and device queries of the C ABI (include/trpx.h).
#include <algorithm>
#include "common.cuh"

namespace trpx {

char* err_buf() {
    static thread_local char buf[512] = "";
    return buf;
}

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace trpx

namespace {
// FP32 FMA peak probe: 16 independent FMA chains per thread, no memory traffic.
__global__ void __launch_bounds__(256) fma_probe_kernel(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    if (s == 12345.678f) out[0] = s;      // never true; keeps the loop alive
}
// KV-window streaming probe: the memory side of the cache-mode rollout WITHOUT its arithmetic.  A warp owns 4
// trajectories; per (step, layer) every lane pulls its share of the key rows, then of the value rows (256-bit loads,
// `rows` of them in flight per lane) of the 4 windows and the warp appends one key and one value row per
// trajectory — the access pattern, granularity and read/write mix of rollout_e32_kernel<4,...>.  Reports the
// achieved DRAM-side rate, i.e. the ceiling of that kernel for this access pattern.
template <int ROWS>
__global__ void __launch_bounds__(1024) window_probe_kernel(float* kv, float* sink, int n_groups, int n_layer, int C,
                                                            int steps, int prefetch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int part = lane & 1, gh = lane >> 1, hh = gh & 3, g = gh >> 2;
    const size_t kv_per_g = (size_t)n_layer * 2 * C * 32;
    float tot = 0.f;
    for (int grp = blockIdx.x * W + warp; grp < n_groups; grp += gridDim.x * W) {
        float* base = kv + (size_t)grp * 4 * kv_per_g;
        for (int s = 0; s < steps; s++) {
            const int slot = s % C;
            for (int l = 0; l < n_layer; l++) {
                float* kl = base + (size_t)(l * 2) * C * 32;
                if (prefetch && lane < 4) {
                    const float* kp = kl + (size_t)lane * kv_per_g;
                    trpx::prefetch_l2_bulk(kp, (uint32_t)C * 128u);
                    trpx::prefetch_l2_bulk(kp + C * 32, (uint32_t)C * 128u);
                }
                for (int gg = 0; gg < 4; gg++) {
                    kl[(size_t)gg * kv_per_g + slot * 32 + lane] = tot;
                    kl[(size_t)gg * kv_per_g + (size_t)C * 32 + slot * 32 + lane] = tot;
                }
                __syncwarp();
                for (int which = 0; which < 2; which++) {
                    const float* Kp = kl + (size_t)g * kv_per_g + (size_t)which * C * 32 + hh * 8;
                    for (int j0 = 0; j0 < C / 2; j0 += ROWS) {
                        float r[ROWS][8];
#pragma unroll
                        for (int jj = 0; jj < ROWS; jj++) trpx::ldg256(Kp + ((j0 + jj) * 2 + part) * 32, r[jj]);
#pragma unroll
                        for (int jj = 0; jj < ROWS; jj++)
#pragma unroll
                            for (int d = 0; d < 8; d++) tot += r[jj][d];
                    }
                }
            }
        }
    }
    if (tot == 12345.678f) sink[0] = tot;
}
// legacy warp-level TF32 MMA peak probe (mma.sync.m16n8k8 -> HMMA.1688.F32.TF32), 8 independent accumulators
__global__ void __launch_bounds__(256) mma_probe_kernel(float* out, int iters) {
    float c[8][4];
#pragma unroll
    for (int t = 0; t < 8; t++)
#pragma unroll
        for (int i = 0; i < 4; i++) c[t][i] = 0.f;
    const unsigned a0 = threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3, b0 = a0 + 11, b1 = a0 + 13;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int t = 0; t < 8; t++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 8; t++)
#pragma unroll
        for (int i = 0; i < 4; i++) s += c[t][i];
    if (s == 12345.678f) out[0] = s;
}
}  // namespace

extern "C" {

int trpx_probe_tf32_mma(int device, float* tflops_out) {
    TRPX_REQUIRE(tflops_out != nullptr, "null output");
    int n = 0;
    TRPX_CUDA(cudaGetDeviceCount(&n));
    TRPX_REQUIRE(device >= 0 && device < n, "device %d out of range", device);
    trpx::DeviceGuard guard(device);
    int sms = 0;
    TRPX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    float* d = nullptr;
    TRPX_CUDA(cudaMalloc(&d, 4));
    cudaEvent_t e0, e1;
    TRPX_CUDA(cudaEventCreate(&e0));
    TRPX_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 2, iters = 20000;
    mma_probe_kernel<<<blocks, 256>>>(d, 100);
    float best = 0.f;
    for (int rep = 0; rep < 3; rep++) {
        TRPX_CUDA(cudaEventRecord(e0));
        mma_probe_kernel<<<blocks, 256>>>(d, iters);
        TRPX_CUDA(cudaEventRecord(e1));
        TRPX_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        TRPX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 16 * 8 * 8 * 8.0 * (double)iters * (double)blocks * 8.0;
        const float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops_out = best;
    return TRPX_OK;
}

int trpx_probe_window_stream(int device, int n_traj, int n_ctx, int n_layer, int steps, int warps_per_cta,
                             int rows_in_flight, int prefetch, float* gbps_out, float* ms_out) {
    TRPX_REQUIRE(gbps_out != nullptr, "null output");
    TRPX_REQUIRE(n_traj > 0 && n_traj % 4 == 0, "n_traj must be a positive multiple of 4");
    TRPX_REQUIRE(n_ctx >= 16 && n_ctx % 16 == 0 && n_layer > 0 && steps > 0, "bad window shape");
    TRPX_REQUIRE(warps_per_cta >= 1 && warps_per_cta <= 32, "warps_per_cta must be 1..32");
    TRPX_REQUIRE(rows_in_flight == 4 || rows_in_flight == 8 || rows_in_flight == 16, "rows_in_flight must be 4, 8 or 16");
    TRPX_REQUIRE((n_ctx / 2) % rows_in_flight == 0, "n_ctx/2 must be a multiple of rows_in_flight");
    int n = 0;
    TRPX_CUDA(cudaGetDeviceCount(&n));
    TRPX_REQUIRE(device >= 0 && device < n, "device %d out of range", device);
    trpx::DeviceGuard guard(device);
    int sms = 0;
    TRPX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int n_groups = n_traj / 4;
    // like the rollout, only the windows of the resident warps exist (grid x warps x 4 trajectories)
    const int grid = std::min(sms, (n_groups + warps_per_cta - 1) / warps_per_cta);
    const size_t kv_per_g = (size_t)n_layer * 2 * n_ctx * 32;
    const size_t resident = (size_t)grid * warps_per_cta * 4;
    float *kv = nullptr, *sink = nullptr;
    TRPX_CUDA(cudaMalloc(&kv, resident * kv_per_g * sizeof(float)));
    TRPX_CUDA(cudaMalloc(&sink, 4));
    TRPX_CUDA(cudaMemset(kv, 0, resident * kv_per_g * sizeof(float)));
    cudaEvent_t e0, e1;
    TRPX_CUDA(cudaEventCreate(&e0));
    TRPX_CUDA(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    // every group reuses the ring of its warp (base index = resident warp), exactly like the rollout workspace
    const int resident_groups = grid * warps_per_cta;
    const int rounds = (n_groups + resident_groups - 1) / resident_groups;
    for (int rep = 0; rep < 3; rep++) {
        TRPX_CUDA(cudaEventRecord(e0));
        for (int r = 0; r < 1; r++) {
            switch (rows_in_flight) {
                case 4: window_probe_kernel<4><<<grid, warps_per_cta * 32>>>(kv, sink, resident_groups, n_layer, n_ctx, steps * rounds, prefetch); break;
                case 8: window_probe_kernel<8><<<grid, warps_per_cta * 32>>>(kv, sink, resident_groups, n_layer, n_ctx, steps * rounds, prefetch); break;
                default: window_probe_kernel<16><<<grid, warps_per_cta * 32>>>(kv, sink, resident_groups, n_layer, n_ctx, steps * rounds, prefetch); break;
            }
        }
        TRPX_CUDA(cudaEventRecord(e1));
        TRPX_CUDA(cudaEventSynchronize(e1));
        TRPX_CUDA(cudaGetLastError());
        float ms = 0.f;
        TRPX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(kv);
    cudaFree(sink);
    // bytes: per (trajectory, step, layer) 2 windows of n_ctx rows of 128 B read + 2 rows written
    const double bytes = (double)resident * (double)(steps * rounds) * n_layer * (2.0 * n_ctx * 128.0 + 2.0 * 128.0);
    *gbps_out = (float)(bytes / (best_ms * 1e-3) / 1e9);
    if (ms_out) *ms_out = best_ms;
    return TRPX_OK;
}

int trpx_probe_fp32_fma(int device, float* tflops_out) {
    TRPX_REQUIRE(tflops_out != nullptr, "null output");
    int n = 0;
    TRPX_CUDA(cudaGetDeviceCount(&n));
    TRPX_REQUIRE(device >= 0 && device < n, "device %d out of range", device);
    trpx::DeviceGuard guard(device);
    int sms = 0;
    TRPX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    float* d = nullptr;
    TRPX_CUDA(cudaMalloc(&d, 4));
    cudaEvent_t e0, e1;
    TRPX_CUDA(cudaEventCreate(&e0));
    TRPX_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 8, iters = 1 << 15;
    fma_probe_kernel<<<blocks, 256>>>(d, 1 << 10, 1.0000001f, 1e-9f);      // warm-up
    float best = 0.f;
    for (int rep = 0; rep < 5; rep++) {
        TRPX_CUDA(cudaEventRecord(e0));
        fma_probe_kernel<<<blocks, 256>>>(d, iters, 1.0000001f, 1e-9f);
        TRPX_CUDA(cudaEventRecord(e1));
        TRPX_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        TRPX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
        const float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops_out = best;
    return TRPX_OK;
}

const char* trpx_last_error(void) { return trpx::err_buf(); }

int trpx_version(void) { return 100; }

int trpx_device_info(int device, int* sm_count, int* smem_optin_bytes) {
    int n = 0;
    TRPX_CUDA(cudaGetDeviceCount(&n));
    TRPX_REQUIRE(device >= 0 && device < n, "device %d out of range (%d devices)", device, n);
    if (sm_count) TRPX_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device));
    if (smem_optin_bytes)
        TRPX_CUDA(cudaDeviceGetAttribute(smem_optin_bytes, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    return TRPX_OK;
}

}  // extern "C"
