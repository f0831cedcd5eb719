// libtrpx: error plumbing, version and device queries of the C ABI (include/trpx.h).
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
