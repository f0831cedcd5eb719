// mma.sync TF32 throughput with distinct operand registers per instruction (no .reuse), like a real GEMV,
// and with the 3xTF32 dependency pattern (3 MMAs per accumulator back to back).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, unsigned seed) {
    float c[4][4];
    for (int t = 0; t < 4; t++) for (int i = 0; i < 4; i++) c[t][i] = 0.f;
    unsigned ah[4][4], al[4][4], b[8];
    for (int s = 0; s < 4; s++) for (int i = 0; i < 4; i++) { ah[s][i] = seed * (s * 4 + i + 1) + threadIdx.x; al[s][i] = ah[s][i] ^ 0x5555; }
    for (int i = 0; i < 8; i++) b[i] = seed + i * 17 + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++)
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (MODE == 0) { mma(c[t], ah[s], b[t], b[t + 4]); mma(c[t], ah[s], b[t], b[t + 4]); mma(c[t], ah[s], b[t], b[t + 4]); }
                else { mma(c[t], al[s], b[t], b[t + 4]); mma(c[t], ah[s], b[(t + 1) & 3], b[4 + ((t + 1) & 3)]); mma(c[t], ah[s], b[t], b[t + 4]); }
            }
        b[it & 7] += 1;
    }
    float sum = 0.f;
    for (int t = 0; t < 4; t++) for (int i = 0; i < 4; i++) sum += c[t][i];
    if (sum == 123.456f) out[0] = sum;
}
template <int MODE> void run(const char* name, int sms, float* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps = 4; wps <= 16; wps *= 2) {
        int blocks = sms * (wps / 8 > 0 ? wps / 8 : 1), threads = wps >= 8 ? 256 : wps * 32, iters = 4000;
        probe<MODE><<<blocks, threads>>>(d, 10, 3);
        cudaEventRecord(e0);
        probe<MODE><<<blocks, threads>>>(d, iters, 3);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 16 * 8 * 8 * 48.0 * iters * (double)blocks * (threads / 32);
        printf("%s warps/SM=%d  %.1f TFLOP/s\n", name, wps, flops / ms / 1e9);
    }
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* d; cudaMalloc(&d, 4);
    run<0>("same-operands x3 chains", sms, d);
    run<1>("3xTF32 pattern (hi/lo operands)", sms, d);
    return 0;
}
