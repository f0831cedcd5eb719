// tcgen05 (UMMA) kind::tf32 probe for sm_100a: validates the descriptor conventions the Cylinder decoder relies on and
// measures what bounds small-N MMAs.  Standalone program (no libtrpx):
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/umma_tf32_probe scripts/probes/umma_tf32_probe.cu
//   scripts/probes/umma_tf32_probe
//
// 1. correctness: D[128 x N] = A[128 x K] . B[N x K]^T with both operands K-major in the NO-SWIZZLE canonical layout
//    (core matrix = 8 rows x 16 bytes stored contiguously; SBO = byte stride between 8-row groups, LBO = byte stride
//    between the two 16-byte K chunks of one K=8 instruction), the A start address shifted by whole rows (the "flat
//    patch" trick of the implicit-GEMM convolution: a 3x3 tap is a row shift of the same shared-memory image).
//    Inputs are small integers, so every product and sum is exact and the comparison is bit-exact.  Both LBO/SBO
//    assignments are tried and reported.
// 2. accumulation accuracy: K = 1152 random TF32-exact inputs against a double reference (mean signed error tells
//    truncation from round-to-nearest accumulation).
// 3. rate: cycles per M128 x N x K8 MMA with A and B from shared memory, for N = 16 ... 256, all 148 SMs busy.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                  // descriptor version 1 (sm_100)
    return d;                                                // base_offset 0, layout_type 0 = no swizzle
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    // c_format F32 (1) @4, a_format TF32 (2) @7, b_format TF32 (2) @10, K-major A and B, N>>3 @17, M>>4 @24
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
#define FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")
#define FENCE_PROXY() asm volatile("fence.proxy.async.shared::cta;" ::: "memory")

// ---------------------------------------------------------------------------------------------------------------
// D[128 x N] = sum over `nchunk` K chunks of A_chunk . B_chunk^T; every chunk is copied global -> shared by the CTA
// (generic proxy), fenced, and consumed by ksteps K=8 MMAs issued by one thread.
__global__ void __launch_bounds__(128) probe_gemm(const float* __restrict__ Aimg, int a_floats, const float* __restrict__ Bimg,
                                                  int b_floats, float* __restrict__ D, int N, int ksteps, int nchunk,
                                                  uint32_t a_off_bytes, uint32_t lbo_a, uint32_t sbo_a,
                                                  uint32_t a_kstep_bytes, uint32_t lbo_b, uint32_t sbo_b,
                                                  uint32_t b_kstep_bytes) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* sA = reinterpret_cast<float*>(smem);
    float* sB = sA + a_floats;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) mbar_init(&bar, 1);
    FENCE_BEFORE();
    __syncthreads();
    FENCE_AFTER();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc(128, N);
    for (int c = 0; c < nchunk; c++) {
        for (int i = tid; i < a_floats; i += 128) sA[i] = Aimg[(size_t)c * a_floats + i];
        for (int i = tid; i < b_floats; i += 128) sB[i] = Bimg[(size_t)c * b_floats + i];
        FENCE_PROXY();
        __syncthreads();
        if (tid == 0) {
            FENCE_AFTER();
            for (int k = 0; k < ksteps; k++) {
                const uint64_t ad = make_desc(smem_u32(sA) + a_off_bytes + k * a_kstep_bytes, lbo_a, sbo_a);
                const uint64_t bd = make_desc(smem_u32(sB) + k * b_kstep_bytes, lbo_b, sbo_b);
                umma_tf32(tmem, ad, bd, idesc, (c > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&bar);
        }
        mbar_wait(&bar, c & 1);          // the chunk's MMAs have read shared memory: it may be overwritten
        FENCE_AFTER();
        __syncthreads();
    }
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 16; j++) D[(size_t)(warp * 32 + lane) * N + c0 + j] = v[j];
    }
    FENCE_BEFORE();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------------
// rate: one thread issues `iters` x 16 MMAs (16 distinct 4 KB A slices, `nb` distinct B slices) into two alternating
// accumulators; cycles from first issue to completion.
__global__ void __launch_bounds__(128) probe_rate(int N, int iters, int same_a, int misalign_rows, unsigned long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* s = reinterpret_cast<float*>(smem);
    for (int i = tid; i < (64 * 1024 + 64 * 1024 + 4096) / 4; i += 128) s[i] = 0.0f;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) mbar_init(&bar, 1);
    FENCE_PROXY();
    FENCE_BEFORE();
    __syncthreads();
    FENCE_AFTER();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N);
        const uint32_t a0 = smem_u32(smem) + misalign_rows * 16, b0 = smem_u32(smem) + 64 * 1024;
        // A slice j: 128 rows x 8 k at 16-byte row pitch: chunk 0 at +0, chunk 1 at +2048 (LBO), SBO = 128
        const unsigned long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint64_t ad = make_desc(a0 + (same_a ? 0 : j * 4096), 2048, 128);
                const uint64_t bd = make_desc(b0 + 4096 + (j & 3) * (uint32_t)(N * 32), (uint32_t)N * 16, 128);
                umma_tf32(tmem + (uint32_t)((j & 1) * 256), ad, bd, idesc, 1u);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const unsigned long long t1 = clock64();
        cyc[blockIdx.x] = t1 - t0;
    }
    FENCE_BEFORE();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
static float tf32_round(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device: %s, %d SMs\n", prop.name, prop.multiProcessorCount);
    CK(cudaFuncSetAttribute(probe_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));

    // ---- 1. correctness ----
    const int Ns[] = {16, 48, 96, 192, 256};
    for (int swap = 0; swap < 1; swap++) {   // (the swapped assignment was tried once: it faults, as it should)
        int all_ok = 1;
        for (int N : Ns) {
            for (int shift : {0, 1, 5, 133}) {
                const int K = 32, NPOS = 128 + 136;
                std::vector<float> A((size_t)NPOS * K), B((size_t)N * K);
                srand(1234 + N + shift);
                for (auto& v : A) v = (float)((rand() % 17) - 8);
                for (auto& v : B) v = (float)((rand() % 9) - 4) * 0.5f;
                std::vector<float> Aimg((size_t)NPOS * K), Bimg((size_t)N * K);
                for (int kc = 0; kc < K / 4; kc++)
                    for (int p = 0; p < NPOS; p++)
                        for (int e = 0; e < 4; e++) Aimg[((size_t)kc * NPOS + p) * 4 + e] = A[(size_t)p * K + kc * 4 + e];
                for (int kc = 0; kc < K / 4; kc++)
                    for (int n = 0; n < N; n++)
                        for (int e = 0; e < 4; e++) Bimg[((size_t)kc * N + n) * 4 + e] = B[(size_t)n * K + kc * 4 + e];
                float *dA, *dB, *dD;
                CK(cudaMalloc(&dA, Aimg.size() * 4));
                CK(cudaMalloc(&dB, Bimg.size() * 4));
                CK(cudaMalloc(&dD, (size_t)128 * N * 4));
                CK(cudaMemcpy(dA, Aimg.data(), Aimg.size() * 4, cudaMemcpyHostToDevice));
                CK(cudaMemcpy(dB, Bimg.data(), Bimg.size() * 4, cudaMemcpyHostToDevice));
                CK(cudaMemset(dD, 0xff, (size_t)128 * N * 4));
                uint32_t lbo_a = NPOS * 16, sbo_a = 128, lbo_b = N * 16, sbo_b = 128;
                if (swap) { std::swap(lbo_a, sbo_a); std::swap(lbo_b, sbo_b); }
                const size_t smem = (Aimg.size() + Bimg.size()) * 4;
                probe_gemm<<<1, 128, smem>>>(dA, (int)Aimg.size(), dB, (int)Bimg.size(), dD, N, K / 8, 1, shift * 16, lbo_a, sbo_a,
                                             2 * NPOS * 16, lbo_b, sbo_b, 2 * N * 16);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
                std::vector<float> D((size_t)128 * N);
                CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
                int bad = 0;
                for (int m = 0; m < 128; m++)
                    for (int n = 0; n < N; n++) {
                        float ref = 0.f;
                        for (int k = 0; k < K; k++) ref += A[(size_t)(m + shift) * K + k] * B[(size_t)n * K + k];
                        if (ref != D[(size_t)m * N + n]) bad++;
                    }
                if (bad) all_ok = 0;
                printf("gemm %s N=%3d shift=%3d: %s (%d mismatches of %d)\n", swap ? "LBO<->SBO swapped" : "LBO=K-chunk stride, SBO=8-row stride",
                       N, shift, bad ? "MISMATCH" : "exact", bad, 128 * N);
                cudaFree(dA); cudaFree(dB); cudaFree(dD);
            }
        }
        printf("=> convention %d %s\n", swap, all_ok ? "CORRECT" : "wrong");
    }

    // ---- 2. accumulation accuracy, K = 1152 in 9 chunks of 128 ----
    {
        const int N = 64, KC = 128, NCH = 9, K = KC * NCH, NPOS = 128;
        std::vector<float> A((size_t)128 * K), B((size_t)N * K);
        srand(77);
        auto rnd = []() { float u = 0; for (int i = 0; i < 12; i++) u += (float)rand() / RAND_MAX; return u - 6.0f; };
        for (auto& v : A) v = tf32_round(rnd());
        for (auto& v : B) v = tf32_round(rnd() * 0.05f);
        std::vector<float> Aimg((size_t)NCH * NPOS * KC), Bimg((size_t)NCH * N * KC);
        for (int c = 0; c < NCH; c++)
            for (int kc = 0; kc < KC / 4; kc++) {
                for (int p = 0; p < NPOS; p++)
                    for (int e = 0; e < 4; e++)
                        Aimg[(size_t)c * NPOS * KC + ((size_t)kc * NPOS + p) * 4 + e] = A[(size_t)p * K + c * KC + kc * 4 + e];
                for (int n = 0; n < N; n++)
                    for (int e = 0; e < 4; e++)
                        Bimg[(size_t)c * N * KC + ((size_t)kc * N + n) * 4 + e] = B[(size_t)n * K + c * KC + kc * 4 + e];
            }
        float *dA, *dB, *dD;
        CK(cudaMalloc(&dA, Aimg.size() * 4));
        CK(cudaMalloc(&dB, Bimg.size() * 4));
        CK(cudaMalloc(&dD, (size_t)128 * N * 4));
        CK(cudaMemcpy(dA, Aimg.data(), Aimg.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, Bimg.data(), Bimg.size() * 4, cudaMemcpyHostToDevice));
        const size_t smem = (size_t)(NPOS * KC + N * KC) * 4;
        probe_gemm<<<1, 128, smem>>>(dA, NPOS * KC, dB, N * KC, dD, N, KC / 8, NCH, 0, NPOS * 16, 128, 2 * NPOS * 16, N * 16, 128, 2 * N * 16);
        CK(cudaDeviceSynchronize());
        std::vector<float> D((size_t)128 * N);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double num = 0, den = 0, bias = 0, absref = 0, f32num = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < N; n++) {
                double ref = 0;
                float f32 = 0.f;
                for (int k = 0; k < K; k++) {
                    ref += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
                    f32 = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], f32);
                }
                const double d = (double)D[(size_t)m * N + n] - ref;
                num += d * d; den += ref * ref; bias += d * (ref >= 0 ? 1 : -1); absref += fabs(ref);
                f32num += ((double)f32 - ref) * ((double)f32 - ref);
            }
        printf("accumulation K=%d: rel-L2 error %.3e (sequential FP32 FMA chain: %.3e); mean signed error toward |larger| / mean |ref| = %.3e\n",
               K, sqrt(num / den), sqrt(f32num / den), bias / absref);
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }

    // ---- 3. rate ----
    {
        unsigned long long* dc;
        const int nb = prop.multiProcessorCount;
        CK(cudaMalloc(&dc, nb * sizeof(unsigned long long)));
        std::vector<unsigned long long> hc(nb);
        for (int mis = 0; mis < 2; mis++)
            for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
                const int same_a = 0, misalign = mis ? 2 : 0;      // A start 2 rows (32 B) off the 128-byte core-matrix boundary
                const int iters = 256;
                probe_rate<<<nb, 128, 136 * 1024>>>(N, iters, same_a, misalign, dc);      // warm-up
                probe_rate<<<nb, 128, 136 * 1024>>>(N, iters, same_a, misalign, dc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("rate kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
                CK(cudaMemcpy(hc.data(), dc, nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                double mean = 0;
                unsigned long long mx = 0;
                for (auto c : hc) { mean += (double)c; mx = std::max(mx, c); }
                mean /= nb;
                const double per = mean / (iters * 16.0);
                printf("rate M128 N=%3d K8 tf32 (%s): %.1f cycles per MMA (max CTA %.1f) -> %.0f MAC/cycle/SM; math floor N/2 = %.0f cycles, "
                       "smem-operand model (4096+32N)/128 = %.0f cycles\n",
                       N, mis ? "A start misaligned by 2 rows" : "A start 128 B aligned", per, (double)mx / (iters * 16.0), 128.0 * N * 8 / per, N / 2.0,
                       (4096 + 32.0 * N) / 128);
            }
        cudaFree(dc);
    }
    return 0;
}
