// K3u: Lorenz / Rossler `recover` for large N on tcgen05 (UMMA) tensor cores.
// Reference: trphysx/embedding/embedding_lorenz.py:107-118 (recover: Linear(32, 500) -> ReLU -> Linear(500, 3) ->
// sd * y + mu), Rossler identical (examples/rossler/rossler_module/embedding_rossler.py:104-115).
//
// Layer 1 ([N x 32] . [32 x 500]) is 91 % of the FLOPs.  A persistent CTA per SM keeps the layer-1 weights as TF32 hi / lo
// planes in shared memory (K-major no-swizzle core-matrix order [k/4][n][4 floats], hidden width padded to 512 with zero
// rows) and runs 128-sample tiles through
//   * 2 producer warps: thread = two sample rows; loads the 32 inputs, splits them into TF32 hi / lo and stores the two
//     A planes ([k/4][row][4 floats]) of a two-stage ring (generic-proxy stores + fence.proxy.async);
//   * 1 MMA warp (one lane): per half of the hidden units (256 columns of tensor memory each, so the epilogue of one half
//     overlaps the MMAs of the other) 12 MMAs M128 x N256 x K8 kind::tf32 = the error-compensated product
//     a_hi.w_hi + a_lo.w_hi + a_hi.w_lo (FP32 parity, as in the Cylinder convs), tcgen05.commit -> mbarrier;
//   * 16 epilogue warps (4 per TMEM lane quarter, 64 columns of a half each): tcgen05.ld -> + bias -> ReLU -> layer 2 in
//     plain FP32 (3 FMAs per hidden unit against one broadcast LDS.128 of {V2[0][u], V2[1][u], V2[2][u], c1[u]}), the
//     four column slices of a row are summed through shared memory, then sd * (y + c2) + mu -> global.
#include "common.cuh"
#include "mlpemb.cuh"
#include "mlpemb_umma.cuh"
#include <algorithm>

using namespace trpx;

namespace {

constexpr int RU_EPI_WARPS = 16, RU_PROD_WARPS = 2;
constexpr int RU_MMA_WARP = RU_EPI_WARPS;
constexpr int RU_PROD_WARP0 = RU_EPI_WARPS + 1;
constexpr int RU_THREADS = (RU_EPI_WARPS + 1 + RU_PROD_WARPS) * 32;       // 608 (<= 104 registers per thread)
constexpr int RU_M = 128;                                                 // samples per tile
constexpr int RU_NP = 512;                                                // hidden width, padded
constexpr int RU_B_PLANE = 8 * RU_NP * 16;                                // bytes of one weight plane [8][512][4 f]
constexpr int RU_A_PLANE = 8 * RU_M * 16;                                 // bytes of one input plane  [8][128][4 f]
constexpr int RU_STAGES = 2;
#ifndef RU_DEBUG
#define RU_DEBUG 0      // timing experiments only (results are garbage): 1 epilogue skips its arithmetic, 2 no MMAs, 4 no input loads
#endif
constexpr size_t RU_SMEM = 2 * (size_t)RU_B_PLANE + RU_STAGES * 2 * (size_t)RU_A_PLANE + RU_NP * 16 /* decB */ +
                           2 * 3 * RU_M * 16 /* partial sums */ + 256 /* barriers, TMEM slot, MSE sums */;

__device__ __forceinline__ uint64_t ru_desc(uint32_t saddr, uint32_t lbo_bytes) {
    // K-major, no swizzle: start >> 4 | LBO >> 4 (stride between the two 16-byte K chunks) | SBO = 128 B (stride between
    // 8-row groups) | descriptor version 1      (same form as cyl_umma.cu, verified by scripts/probes/umma_tf32_probe.cu)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
constexpr uint32_t ru_idesc(int M, int N) {
    // D = F32 (1 @4), A = B = TF32 (2 @7, 2 @10), both K-major, N >> 3 @17, M >> 4 @24
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void ru_mma(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ru_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ru_tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void ru_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void ru_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ru_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ru_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ru_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ru_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ru_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// packed FP32 (f32x2), element-wise: (d0, d1) += (a0, a1) * (b0, b1)   and   (d0, d1) = (a0, a1) + (b0, b1)
__device__ __forceinline__ void ru_ffma2v(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n.reg .b64 ra, rb, rc;\n"
        "mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%0, %1};\n"
        "fma.rn.f32x2 rc, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rc;\n}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void ru_fadd2v(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n.reg .b64 ra, rb, rc;\n"
        "mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n"
        "add.rn.f32x2 rc, ra, rb;\n"
        "mov.b64 {%0, %1}, rc;\n}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// 16 accumulator columns (hidden units u .. u + 15, u even) of a sample row through bias, ReLU and layer 2: two hidden
// units per packed instruction, even / odd partial sums (y[0..1] output 0, y[2..3] output 1, y[4..5] output 2)
__device__ __forceinline__ void ru_layer2_16(const uint32_t (&r)[16], const float4* __restrict__ tab, float (&y)[6]) {
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
        const float4 wa = tab[c], wb = tab[c + 1];           // {w0a, w0b, w1a, w1b}, {w2a, w2b, ba, bb}
        float va, vb;
        ru_fadd2v(va, vb, __uint_as_float(r[c]), __uint_as_float(r[c + 1]), wb.z, wb.w);
        va = fmaxf(va, 0.f);
        vb = fmaxf(vb, 0.f);
        ru_ffma2v(y[0], y[1], va, vb, wa.x, wa.y);
        ru_ffma2v(y[2], y[3], va, vb, wa.z, wa.w);
        ru_ffma2v(y[4], y[5], va, vb, wb.x, wb.y);
    }
}
__device__ __forceinline__ void ru_split4(const float4 v, float4& hi, float4& lo) {
    hi.x = __uint_as_float(tf32_rna_bits(v.x)); lo.x = v.x - hi.x;
    hi.y = __uint_as_float(tf32_rna_bits(v.y)); lo.y = v.y - hi.y;
    hi.z = __uint_as_float(tf32_rna_bits(v.z)); lo.z = v.z - hi.z;
    hi.w = __uint_as_float(tf32_rna_bits(v.w)); lo.w = v.w - hi.w;
}

__global__ void __launch_bounds__(RU_THREADS, 1) mlp_recover_umma_kernel(const float* __restrict__ blob, MlpLayout L,
                                                                         const float* __restrict__ g,
                                                                         float* __restrict__ x, long long N,
                                                                         const float* __restrict__ target,
                                                                         float* __restrict__ partial, int debug) {
    extern __shared__ __align__(128) unsigned char ru_smem[];
    unsigned char* b_hi = ru_smem;
    unsigned char* b_lo = b_hi + RU_B_PLANE;
    unsigned char* a_st = b_lo + RU_B_PLANE;                              // [stage][hi | lo][8][128][4 f]
    float4* decB = reinterpret_cast<float4*>(a_st + RU_STAGES * 2 * RU_A_PLANE);   // [512] {V2[0][u], V2[1][u], V2[2][u], c1[u]}
    float4* ypart = decB + RU_NP;                                         // [tile parity][slice 1..3][row]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ypart + 2 * 3 * RU_M);
    uint64_t* a_full = bars;                 // [stage]  producers -> MMA           (one arrival per producer warp)
    uint64_t* a_empty = bars + 2;            // [stage]  tcgen05.commit -> producers
    uint64_t* tfull = bars + 4;              // [half]   tcgen05.commit -> epilogue
    uint64_t* tempty = bars + 6;             // [half]   epilogue -> MMA            (16 arrivals: one per epilogue warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    float* red = reinterpret_cast<float*>(bars + 9);       // [4] squared-error sums of the finishing warps (fused recover + MSE)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Hd = L.Hd;
    const long long ntiles = (N + RU_M - 1) / RU_M;

    if (tid == 0) {
        for (int s = 0; s < RU_STAGES; s++) { mbar_init(&a_full[s], RU_PROD_WARPS); mbar_init(&a_empty[s], 1); }
        for (int hf = 0; hf < 2; hf++) { mbar_init(&tfull[hf], 1); mbar_init(&tempty[hf], RU_EPI_WARPS); }
    }
    if (warp == RU_MMA_WARP) ru_tmem_alloc(tmem_slot, 512);
    // layer-1 weights -> TF32 hi / lo planes [k/4][n][4]; hidden units >= Hd are zero rows
    for (int i = tid; i < 8 * RU_NP; i += RU_THREADS) {
        const int kc = i / RU_NP, n = i - kc * RU_NP;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < Hd) v = __ldg(reinterpret_cast<const float4*>(blob + L.decA + (size_t)n * ME + kc * 4));
        float4 hi, lo;
        ru_split4(v, hi, lo);
        reinterpret_cast<float4*>(b_hi)[i] = hi;
        reinterpret_cast<float4*>(b_lo)[i] = lo;
    }
    // layer-2 table in pairs of hidden units (a = 2p, b = 2p + 1): decB[2p] = {w0a, w0b, w1a, w1b}, decB[2p + 1] = {w2a, w2b, c1a, c1b}
    for (int p2 = tid; p2 < RU_NP / 2; p2 += RU_THREADS) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 ua = 2 * p2 < Hd ? __ldg(reinterpret_cast<const float4*>(blob + L.decB) + 2 * p2) : z;
        const float4 ub = 2 * p2 + 1 < Hd ? __ldg(reinterpret_cast<const float4*>(blob + L.decB) + 2 * p2 + 1) : z;
        decB[2 * p2] = make_float4(ua.x, ub.x, ua.y, ub.y);
        decB[2 * p2 + 1] = make_float4(ua.z, ub.z, ua.w, ub.w);
    }
    ru_fence_proxy_async();
    ru_fence_before();
    __syncthreads();
    ru_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < RU_EPI_WARPS) {
        // =============================== epilogue ===============================
        const int q = warp & 3, e = warp >> 2;                 // TMEM lane quarter, 64-column slice of a half
        const int row = q * 32 + lane;                         // sample row of the tile
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float c20 = __ldg(blob + L.c2), c21 = __ldg(blob + L.c2 + 1), c22 = __ldg(blob + L.c2 + 2);
        const float mu0 = __ldg(blob + L.mu), mu1 = __ldg(blob + L.mu + 1), mu2 = __ldg(blob + L.mu + 2);
        const float sd0 = __ldg(blob + L.sd), sd1 = __ldg(blob + L.sd + 1), sd2 = __ldg(blob + L.sd + 2);
        int it = 0;
        float sq = 0.f;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            float y[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int hf = 0; hf < 2; hf++) {
                mbar_wait_backoff(&tfull[hf], it & 1, 32);
                ru_fence_after();
                const int u0 = hf * 256 + e * 64;
                const uint32_t tcol = tmem_base + lane_base + (uint32_t)u0;
                // 16 columns at a time, the next tcgen05.ld in flight under the arithmetic of the previous one; the
                // accumulator half goes back to the MMA warp as soon as its last columns are in registers
                uint32_t ra[16], rb[16];
                ru_tmem_ld16(tcol, ra);
                ru_wait_ld();
                ru_tmem_ld16(tcol + 16, rb);
                if (!(debug & 1)) ru_layer2_16(ra, decB + u0, y);
                ru_wait_ld();
                ru_tmem_ld16(tcol + 32, ra);
                if (!(debug & 1)) ru_layer2_16(rb, decB + u0 + 16, y);
                ru_wait_ld();
                ru_tmem_ld16(tcol + 48, rb);
                if (!(debug & 1)) ru_layer2_16(ra, decB + u0 + 32, y);
                ru_wait_ld();
                ru_fence_before();
                __syncwarp();
                if (lane == 0) ru_arrive(&tempty[hf]);
                if (!(debug & 1)) ru_layer2_16(rb, decB + u0 + 48, y);
            }
            float y0 = y[0] + y[1], y1 = y[2] + y[3], y2 = y[4] + y[5];
            // the four column slices of a row: slices 1..3 go through shared memory, slice 0 finishes the row
            float4* yp = ypart + (it & 1) * 3 * RU_M;
            if (e > 0) yp[(e - 1) * RU_M + row] = make_float4(y0, y1, y2, 0.f);
            asm volatile("bar.sync 1, %0;" ::"n"(RU_EPI_WARPS * 32) : "memory");
            if (e == 0) {
                const float4 p1 = yp[row], p2 = yp[RU_M + row], p3 = yp[2 * RU_M + row];
                y0 = (y0 + p1.x) + (p2.x + p3.x);
                y1 = (y1 + p1.y) + (p2.y + p3.y);
                y2 = (y2 + p1.z) + (p2.z + p3.z);
                const long long i = tile * RU_M + row;
                if (i < N) {
                    const float x0 = sd0 * (y0 + c20) + mu0, x1 = sd1 * (y1 + c21) + mu1, x2 = sd2 * (y2 + c22) + mu2;
                    if (x != nullptr) {
                        x[i * MS + 0] = x0;
                        x[i * MS + 1] = x1;
                        x[i * MS + 2] = x2;
                    }
                    if (target != nullptr) {
                        const float d0 = x0 - __ldg(target + i * MS), d1 = x1 - __ldg(target + i * MS + 1),
                                    d2 = x2 - __ldg(target + i * MS + 2);
                        sq += d0 * d0 + d1 * d1 + d2 * d2;
                    }
                }
            }
        }
        if (target != nullptr && e == 0) {
            sq = warp_sum(sq);
            if (lane == 0) red[q] = sq;
        }
    } else if (warp == RU_MMA_WARP) {
        // =============================== MMA issue ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = ru_idesc(128, 256);
            int it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
                const int st = it & 1;
                mbar_wait(&a_full[st], (it >> 1) & 1);
                ru_fence_after();
                const uint32_t sa_hi = smem_u32(a_st + (size_t)st * 2 * RU_A_PLANE), sa_lo = sa_hi + RU_A_PLANE;
                const uint32_t sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
                for (int hf = 0; hf < 2; hf++) {
                    mbar_wait(&tempty[hf], (it & 1) ^ 1);
                    ru_fence_after();
                    const uint32_t tc = tmem_base + (uint32_t)(hf * 256);
#pragma unroll
                    for (int ks = (debug & 2) ? 4 : 0; ks < 4; ks++) {   // k = 8 ks .. 8 ks + 7: chunks 2 ks, 2 ks + 1
                        const uint64_t ah = ru_desc(sa_hi + ks * 2 * (RU_M * 16), RU_M * 16);
                        const uint64_t al = ru_desc(sa_lo + ks * 2 * (RU_M * 16), RU_M * 16);
                        const uint64_t bh = ru_desc(sb_hi + ks * 2 * (RU_NP * 16) + hf * 256 * 16, RU_NP * 16);
                        const uint64_t bl = ru_desc(sb_lo + ks * 2 * (RU_NP * 16) + hf * 256 * 16, RU_NP * 16);
                        ru_mma(tc, al, bh, idesc, ks > 0 ? 1u : 0u);     // small terms first, the dominant one last
                        ru_mma(tc, ah, bl, idesc, 1u);
                        ru_mma(tc, ah, bh, idesc, 1u);
                    }
                    ru_commit(&tfull[hf]);
                }
                ru_commit(&a_empty[st]);
            }
        }
    } else {
        // =============================== producers ===============================
        const int row0 = (warp - RU_PROD_WARP0) * 32 + lane;            // rows row0 and row0 + 64 of the tile
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int st = it & 1;
            float4* ah = reinterpret_cast<float4*>(a_st + (size_t)st * 2 * RU_A_PLANE);
            float4* al = ah + RU_A_PLANE / 16;
#pragma unroll 1
            for (int hr = 0; hr < 2; hr++) {
                const int row = row0 + 64 * hr;
                const long long i = tile * RU_M + row;
                float4 v[8];
#pragma unroll
                for (int kc = 0; kc < 8; kc++)
                    v[kc] = (i < N && !(debug & 4)) ? __ldg(reinterpret_cast<const float4*>(g + i * ME) + kc) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (hr == 0) mbar_wait_backoff(&a_empty[st], ((it >> 1) & 1) ^ 1, 32);
#pragma unroll
                for (int kc = 0; kc < 8; kc++) {
                    float4 hi, lo;
                    ru_split4(v[kc], hi, lo);
                    ah[kc * RU_M + row] = hi;
                    al[kc * RU_M + row] = lo;
                }
            }
            ru_fence_proxy_async();
            __syncwarp();
            if (lane == 0) ru_arrive(&a_full[st]);
        }
    }
    ru_fence_before();
    __syncthreads();
    if (warp == RU_MMA_WARP) {
        ru_fence_after();
        ru_tmem_dealloc(tmem_base, 512);
    }
    if (target != nullptr && tid == 0) partial[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);     // fixed order
}

}  // namespace

namespace trpx {

bool mlp_recover_umma_ok(trpx_mlpemb_t h, long long N) {
    return h->cfg.hidden <= RU_NP && N >= 4LL * RU_M * h->sm_count && RU_SMEM + 1024 <= (size_t)h->smem_optin;
}

int mlp_recover_umma_grid(trpx_mlpemb_t h, long long N) {
    return (int)std::min<long long>(h->sm_count, (N + RU_M - 1) / RU_M);
}

// target != null: fused recover + MSE (x may then be null); per-CTA squared-error sums go to partial[0 .. grid)
int mlp_recover_umma(trpx_mlpemb_t h, const float* g, float* x, long long N, const float* target, float* partial,
                     cudaStream_t st) {
    const MlpLayout L = MlpLayout::make(h->cfg.hidden);
    TRPX_CUDA(cudaFuncSetAttribute(mlp_recover_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RU_SMEM));
    mlp_recover_umma_kernel<<<mlp_recover_umma_grid(h, N), RU_THREADS, RU_SMEM, st>>>(h->blob, L, g, x, N, target, partial, RU_DEBUG);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

}  // namespace trpx
