// Packed FP32 FMA probe (fma.rn.f32x2 -> SASS FFMA2) against scalar FFMA: nvcc -arch=sm_100a -O3 ffma2_probe.cu && ./a.out
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b) {
    asm("{\n.reg .b64 ra, rb, rc;\n"
        "mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %4};\n mov.b64 rc, {%0, %1};\n"
        "fma.rn.f32x2 rc, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rc;\n}"
        : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1), "f"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; i++) x[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 32; i++) x[i] = fmaf(x[i], a, b);
        } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) ffma2(x[i], x[i + 1], x[i], x[i + 1], a);   // x = x*a + x
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i++) s += x[i];
    if (s == 12345.678f) out[0] = s;
}
int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* d;
    cudaMalloc(&d, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 1 << 14;
    for (int mode = 0; mode < 2; mode++)
        for (int wps = 1; wps <= 8; wps *= 2) {          // CTAs of 256 threads per SM
            const int blocks = sms * wps;
            float best = 0;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) probe<0><<<blocks, 256>>>(d, iters, 1.0000001f, 1e-9f);
                else probe<1><<<blocks, 256>>>(d, iters, 0.5f, 0.f);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                const double tf = 2.0 * 32 * (double)iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
                if (tf > best) best = tf;
            }
            printf("%s  %d CTAs/SM: %.1f TFLOP/s\n", mode ? "FFMA2 (fma.rn.f32x2)" : "FFMA  (fmaf)        ", wps, best);
        }
    return 0;
}
