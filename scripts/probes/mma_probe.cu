// Microbenchmark: legacy mma.sync m16n8k8 TF32 throughput on sm_100a (does the HMMA path pay for 3xTF32?)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) mma_probe(float* out, int iters) {
    float c[8][4];
    for (int t = 0; t < 8; t++) for (int i = 0; i < 4; i++) c[t][i] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3, b0 = a0 + 11, b1 = a0 + 13;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int t = 0; t < 8; t++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
    for (int t = 0; t < 8; t++) for (int i = 0; i < 4; i++) s += c[t][i];
    if (s == 123.456f) out[0] = s;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps = 4; wps <= 32; wps *= 2) {
        int blocks = sms * (wps / 8 > 0 ? wps / 8 : 1), threads = wps >= 8 ? 256 : wps * 32;
        int iters = 20000;
        mma_probe<<<blocks, threads>>>(d, 100);
        cudaEventRecord(e0);
        mma_probe<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * (double)blocks * (threads / 32);
        printf("warps/SM=%d  %.1f TFLOP/s (tf32 mma.sync m16n8k8)  %s\n", wps, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
