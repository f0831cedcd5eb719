// K1t — autoregressive rollout with the KV window kept ON CHIP, in tensor memory (sm_100a).
//
// Replaces the decode loop of trphysx/transformer/generate_utils.py:141-161 with the cache concat of
// attention.py:182-188 for the E = 32 / 4-head / <= 4-layer models (Lorenz, Rossler).  BASELINE north_star: "keeps
// every layer's weights resident in shared memory and the KV cache for n_ctx <= 64 in shared memory or registers".
//
// Placement.  The model's weights (203 KB FP32) fill shared memory, so the windows go where nothing else lives: the
// 256 KB of TENSOR MEMORY of the SM (128 lanes x 512 columns x 32 bit), used here as a lane-private scratchpad through
// tcgen05.st / tcgen05.ld (SASS STTM / LDTM) — no tcgen05.mma touches it.  A CTA (one per SM, persistent) keeps
// 256 / n_ctx trajectories resident for the whole horizon (8 for n_ctx = 32, 4 for n_ctx = 64):
//
//     lane   = (key part kp, trajectory, head)         parts = n_ctx / 8, 128 / parts (trajectory, head) pairs
//     column = layer * 128 + slot_in_part * 16 + {k[0..8), v[0..8)}            (8 slots per lane and layer)
//
// so ring slot `s mod n_ctx` of a (trajectory, head) pair lives in lane part (slot >> 3), columns (slot & 7) * 16.
// Warp w can only reach TMEM lanes 32 (w & 3) .. +32; warps w and w + 4 share a lane quarter and split its 8 slots
// 4 / 4, so every (lane, column) cell is read and written by exactly ONE warp: no cross-warp TMEM hazards, no fences.
// Per layer a thread loads its 4 (k, v) rows with two tcgen05.ld.32x32b.x32, computes a partial softmax (running
// max / sum / weighted values) and the 8 or 16 partials of a pair are merged through shared memory.
//
// The dense part runs on the warp-level tensor-core MMA with the error-compensated 3xTF32 product of the no-cache
// kernel (gpt2.cu, rollout_e32v4_kernel), the 8 warps of the CTA sharing ONE tile of trajectories (rows 8..15 of the
// m16 tile are zero): n-tiles of the QKV and MLP-in products are dealt over the warps, the MLP-out product is split
// over K (each warp consumes the 16 hidden units it just produced, straight from registers) and reduced through
// shared memory; the small output projection is computed redundantly by every warp so that LayerNorm 2 needs no
// barrier.  Five CTA barriers per layer, one per step for the head.
//
// HBM traffic is the algorithmic minimum: 128 B written per trajectory-step, x0 read once.
//
// The LayerNorm of a residual row is applied once, by the phase that owns the row (the projection and the partial-sum
// reduction run one warp per row), not by every consumer.  CACHE = false is the reference's default generate() mode
// (no window, position 0) on the same tile without the window phases.
#include "gpt2_tmem.cuh"
#include <cstdlib>

namespace trpx {
namespace {

constexpr int TT = 256;                   // threads per CTA
constexpr int LDH = 36, LDQ = 132, LDR = 40;     // LDQ: q | k | v (K half 0) | v (K half 1) + 4
constexpr int SCR_FLOATS = 3 * 8 * LDH + 8 * LDQ + 8 * 10 * 32 + 2 * 8 * LDR + 2 * 8 * LDH;
constexpr float QSCALE = 1.4426950408889634f * 0.35355339059327373f;   // log2(e) / sqrt(8)

// Exact two-term split x = hi + lo with hi = x truncated to TF32 (one LOP3 + one FADD; cvt.rna.tf32 is a three-
// instruction sequence on sm_100a).  Truncation instead of rounding is harmless here because the product below keeps
// all four terms.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// A fragment of one k-step for a tile of <= 8 trajectories: rows 0..7 of the m16 tile carry the hi parts, rows 8..15
// the lo parts of the SAME activations (a0/a2: row g = hi, a1/a3: row g+8 = lo; logical k = 8s+2q, 8s+2q+1).
struct AFrag { uint32_t h0, l0, h2, l2; };
__device__ __forceinline__ AFrag make_a(float x0, float x1) {
    AFrag a;
    split_tf32(x0, a.h0, a.l0);
    split_tf32(x1, a.h2, a.l2);
    return a;
}
__device__ __forceinline__ void mma16(float (&c)[4], const AFrag& a, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a.h0), "r"(a.l0), "r"(a.h2), "r"(a.l2), "r"(b0), "r"(b1));
}
// swizzled weight element pair of the B fragment (see gpt2.cu: swizzle_weights_kernel / ldb)
__device__ __forceinline__ void ldb(const float* W, int N, int s, int t, int g, int q, float& w0, float& w1) {
    const float* p = W + (8 * s + 2 * q) * N + 8 * (t ^ q) + g;
    w0 = p[0];
    w1 = p[N];
}
// Error-compensated product in TWO MMAs: ch += [a_hi; a_lo] * w_hi, cl += [a_hi; a_lo] * w_lo.  Rows g of ch / cl hold
// a_hi*w_hi / a_hi*w_lo, rows g+8 (the same lane's c[2], c[3]) a_lo*w_hi / a_lo*w_lo: all four terms of the FP32 product.
__device__ __forceinline__ void mma2(float (&ch)[4], float (&cl)[4], const AFrag& a, float w0, float w1) {
    uint32_t bh0, bl0, bh1, bl1;
    split_tf32(w0, bh0, bl0);
    split_tf32(w1, bh1, bl1);
    mma16(cl, a, bl0, bl1);
    mma16(ch, a, bh0, bh1);
}
// the same with fresh accumulators (D = A*B + 0): independent of every other MMA, so a burst of them pipelines
__device__ __forceinline__ void mma16z(float (&d)[4], const AFrag& a, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a.h0), "r"(a.l0), "r"(a.h2), "r"(a.l2), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void mma2z(float (&dh)[4], float (&dl)[4], const AFrag& a, float w0, float w1) {
    uint32_t bh0, bl0, bh1, bl1;
    split_tf32(w0, bh0, bl0);
    split_tf32(w1, bh1, bl1);
    mma16z(dl, a, bl0, bl1);
    mma16z(dh, a, bh0, bh1);
}
__device__ __forceinline__ float2 fold(const float (&ch)[4], const float (&cl)[4]) {
    return make_float2(ch[0] + (ch[2] + cl[0] + cl[2]), ch[1] + (ch[3] + cl[1] + cl[3]));
}
// sum of two / four folded accumulator pairs: the large (hi*hi) terms last
__device__ __forceinline__ float2 fold2(const float (&h0)[4], const float (&l0)[4], const float (&h1)[4], const float (&l1)[4]) {
    const float sx = (h0[2] + l0[0] + l0[2]) + (h1[2] + l1[0] + l1[2]);
    const float sy = (h0[3] + l0[1] + l0[3]) + (h1[3] + l1[1] + l1[3]);
    return make_float2((h0[0] + h1[0]) + sx, (h0[1] + h1[1]) + sy);
}
// LayerNorm of row g from the lane's 8 values x[2s], x[2s+1] = X[g][8s+2q], X[g][8s+2q+1]; result as A fragments
__device__ __forceinline__ void ln_frags(const float (&x)[8], const float* lw, const float* lb, float eps, int q,
                                         AFrag (&a)[4]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    s += __shfl_xor_sync(FULL, s, 1);
    s += __shfl_xor_sync(FULL, s, 2);
    const float m = s * (1.0f / 32.0f);
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { const float d = x[i] - m; v = fmaf(d, d, v); }
    v += __shfl_xor_sync(FULL, v, 1);
    v += __shfl_xor_sync(FULL, v, 2);
    const float r = 1.0f / sqrtf(v * (1.0f / 32.0f) + eps);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float2 w2 = *reinterpret_cast<const float2*>(lw + 8 * k + 2 * q);
        const float2 b2 = *reinterpret_cast<const float2*>(lb + 8 * k + 2 * q);
        a[k] = make_a((x[2 * k] - m) * r * w2.x + b2.x, (x[2 * k + 1] - m) * r * w2.y + b2.y);
    }
}
// LayerNorm of one row by the warp that owns it (lane = channel): used where a phase produces a full row anyway, so that
// no consumer has to recompute the statistics (the first version normalised redundantly in every warp: ~300 cycles per use)
__device__ __forceinline__ float ln_row(float x, const float* lw, const float* lb, float eps, int lane) {
    // sum and sum of squares travel through the five xor-shuffle steps together (two independent chains instead of two
    // dependent reductions), and the reciprocal standard deviation is one MUFU.RSQ: on this latency-bound path the
    // two-pass / IEEE sqrt + divide form cost ~500 cycles per use.  Residual-stream rows are O(1) with |mean| <~ std,
    // so E[x^2] - mean^2 loses nothing that matters (parity tests unchanged at ~8e-7 against the 1e-5 bar).
    float s1 = x, s2 = x * x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(FULL, s1, o);
        s2 += __shfl_xor_sync(FULL, s2, o);
    }
    const float mean = s1 * (1.0f / 32.0f);
    const float var = fmaxf(fmaf(-mean, mean, s2 * (1.0f / 32.0f)), 0.f);
    return (x - mean) * rsqrtf(var + eps) * lw[lane] + lb[lane];
}
__device__ __forceinline__ void frags_from_rows(const float* X, int ld, int g, int q, AFrag (&a)[4]) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float2 v = *reinterpret_cast<const float2*>(X + g * ld + 8 * k + 2 * q);
        a[k] = make_a(v.x, v.y);
    }
}
__device__ __forceinline__ void load_row_frag(const float* X, int ld, int g, int q, float (&x)[8]) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float2 v = *reinterpret_cast<const float2*>(X + g * ld + 8 * k + 2 * q);
        x[2 * k] = v.x;
        x[2 * k + 1] = v.y;
    }
}

// ---- tensor memory as a scratchpad ----
__device__ __forceinline__ void tmem_alloc512(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512u) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
        "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// PROF: thread 0 of CTA 0 accumulates the cycles between the phase boundaries of the layer loop into prof[0..8)
// (debugging aid, TRPX_TMEM_PROF=1)
#define TMEM_TICK(i)                                                     \
    if constexpr (PROF) {                                                \
        if (tid == 0) {                                                  \
            const long long now = clock64();                             \
            pacc[i] += now - tprev;                                      \
            tprev = now;                                                 \
        }                                                                \
    }

// CACHE = false is the reference's DEFAULT generate() mode (use_cache=False: a one-token forward at position 0, the head
// output is v itself): the same cooperative tile without the window phases, for the latency-bound small batches.
template <int NCTX, bool PROF, bool CACHE = true>
__global__ void __launch_bounds__(TT, 1) rollout_e32t_kernel(TmemRolloutParams p, long long* prof) {
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = 0;
    constexpr int TRAJ = 256 / NCTX;          // trajectories resident per CTA
    constexpr int PAIRS = TRAJ * 4;           // (trajectory, head) pairs = lanes per key part
    constexpr int SUB = 32 / PAIRS;           // key parts inside one warp's lane quarter
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const Gpt2Layout& L = p.lay;
    const int NL = L.n_layer;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    float* wsm = sm;
    float* sH = sm + L.total;                 // residual stream [8][LDH]
    float* sHn = sH + 8 * LDH;                // next step's (double buffer across the head)
    float* sAO = sHn + 8 * LDH;               // merged attention output [8][LDH]
    float* sQKV = sAO + 8 * LDH;              // q | k | v rows [8][LDQ]
    float* sP = sQKV + 8 * LDQ;               // attention partials [8 warps][10][32] / MLP-out partials [8][8][LDR]
    float* sHm = sP + 8 * 10 * 32;            // residual after the output projection [8][LDH]
    float* sA1 = sHm + 2 * 8 * LDR;           // LayerNorm-1 (or final LayerNorm) output of the current residual [8][LDH]
    float* sA2 = sA1 + 8 * LDH;               // LayerNorm-2 output [8][LDH]

    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) tmem_alloc512(&tmem_slot);
    for (int i = tid; i < SCR_FLOATS; i += TT) sH[i] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)L.total * 4u;
        mbar_expect_tx(&bar, bytes);
        const uint32_t CH = 32768u;
        for (uint32_t off = 0; off < bytes; off += CH)
            bulk_g2s(reinterpret_cast<char*>(wsm) + off, reinterpret_cast<const char*>(p.blob_sw) + off,
                     min(CH, bytes - off), &bar);
    }
    const uint32_t tlane = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    mbar_wait(&bar, 0);

    // attention role of the thread
    const int pair = lane % PAIRS, sub = lane / PAIRS;
    const int atraj = pair >> 2, ahead = pair & 3;
    const int kpart = (warp & 3) * SUB + sub, half = warp >> 2;
    const int sbase = kpart * 8 + half * 4;               // first ring slot of the thread's 4
    const float* gw = wsm + L.lnfw;

    const int ngroups = (p.B + TRAJ - 1) / TRAJ;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        // ---- group start: clear the warp's window cells, h = x0 + PE(0) ----
        if constexpr (CACHE) {
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; i++) z[i] = 0u;
            for (int l = 0; l < NL; l++)
#pragma unroll
                for (int c = 0; c < 4; c++) tmem_st16(tlane + l * 128 + half * 64 + c * 16, z);
            tmem_wait_st();
        }
        {
            const int r = tid >> 5, c = tid & 31;
            const int b = grp * TRAJ + r;
            float v = 0.f;
            if (r < TRAJ && b < p.B) v = p.x0[(size_t)b * p.x0_stride + c] + __ldg(p.pe + c);
            sH[r * LDH + c] = v;
            sA1[r * LDH + c] = ln_row(v, wsm + L.ln1w, wsm + L.ln1b, p.eps, c);       // LayerNorm 1 of layer 0
        }
        __syncthreads();

        if constexpr (PROF) tprev = clock64();
        for (int s = 0; s < p.S; s++) {
            const int slot = s & (NCTX - 1), kp_c = slot >> 3, j_c = slot & 7;
            const bool full = (s + 1 >= NCTX);
            for (int l = 0; l < NL; l++) {
                const float* w = wsm + l * L.layer_stride;
                // ---- phase A: LN1 + this warp's share of the QKV product: one full n-tile (q tiles on warps 0..3, k
                //      tiles on warps 4..7) and HALF the K range of a v tile (two partial v rows, summed by the reader) ----
                {
                    AFrag a[4];
                    frags_from_rows(sA1, LDH, g, q, a);                 // LayerNorm 1 was applied by the row owner
                    const int tv = 8 + (warp & 3), kv = 2 * (warp >> 2);
                    float dh[4][4], dl[4][4], eh[2][4], el[2][4];      // one accumulator pair per k-step: no MMA chains
                    const AFrag av[2] = {warp < 4 ? a[0] : a[2], warp < 4 ? a[1] : a[3]};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float w0, w1;
                        if constexpr (CACHE) {
                            ldb(w + L.wqkv, 96, k, warp, g, q, w0, w1);
                            mma2z(dh[k], dl[k], a[k], w0, w1);
                        }
                        if (k < 2) {
                            ldb(w + L.wqkv, 96, kv + k, tv, g, q, w0, w1);
                            mma2z(eh[k], el[k], av[k], w0, w1);
                        }
                    }
                    if constexpr (CACHE) {
                        const float2 b2 = *reinterpret_cast<const float2*>(w + L.bqkv + 8 * warp + 2 * q);
                        float2 y;
                        {
                            const float2 y0 = fold2(dh[0], dl[0], dh[1], dl[1]), y1 = fold2(dh[2], dl[2], dh[3], dl[3]);
                            y.x = y0.x + y1.x + b2.x;
                            y.y = y0.y + y1.y + b2.y;
                        }
                        if (warp < 4) { y.x *= QSCALE; y.y *= QSCALE; }
                        *reinterpret_cast<float2*>(sQKV + g * LDQ + 8 * warp + 2 * q) = y;
                    }
                    float2 v = fold2(eh[0], el[0], eh[1], el[1]);
                    if (warp < 4) {
                        const float2 bv = *reinterpret_cast<const float2*>(w + L.bqkv + 8 * tv + 2 * q);
                        v.x += bv.x; v.y += bv.y;
                    }
                    *reinterpret_cast<float2*>(sQKV + g * LDQ + 8 * tv + 32 * (warp >> 2) + 2 * q) = v;
                }
                TMEM_TICK(0)
                __syncthreads();
                TMEM_TICK(1)
                if constexpr (CACHE) {
                // ---- phase B: append the new (k, v) row, partial softmax over the thread's 4 slots ----
                {
                    const float* qrow = sQKV + atraj * LDQ + ahead * 8;
                    const uint32_t tcol = tlane + l * 128;
                    if (warp == (kp_c / SUB) + 4 * (j_c >> 2)) {          // warp-uniform: this warp owns the slot
                        uint32_t nv[16];
                        const float4 k0 = *reinterpret_cast<const float4*>(qrow + 32), k1 = *reinterpret_cast<const float4*>(qrow + 36);
                        float4 v0 = *reinterpret_cast<const float4*>(qrow + 64), v1 = *reinterpret_cast<const float4*>(qrow + 68);
                        {
                            const float4 u0 = *reinterpret_cast<const float4*>(qrow + 96), u1 = *reinterpret_cast<const float4*>(qrow + 100);
                            v0.x += u0.x; v0.y += u0.y; v0.z += u0.z; v0.w += u0.w;
                            v1.x += u1.x; v1.y += u1.y; v1.z += u1.z; v1.w += u1.w;
                        }
                        nv[0] = __float_as_uint(k0.x); nv[1] = __float_as_uint(k0.y); nv[2] = __float_as_uint(k0.z); nv[3] = __float_as_uint(k0.w);
                        nv[4] = __float_as_uint(k1.x); nv[5] = __float_as_uint(k1.y); nv[6] = __float_as_uint(k1.z); nv[7] = __float_as_uint(k1.w);
                        nv[8] = __float_as_uint(v0.x); nv[9] = __float_as_uint(v0.y); nv[10] = __float_as_uint(v0.z); nv[11] = __float_as_uint(v0.w);
                        nv[12] = __float_as_uint(v1.x); nv[13] = __float_as_uint(v1.y); nv[14] = __float_as_uint(v1.z); nv[15] = __float_as_uint(v1.w);
                        if constexpr (SUB > 1) {
                            // the slot belongs to one key part = half of the warp's lanes; the others write back
                            // what they hold (tcgen05.st is warp-wide)
                            uint32_t old[16];
                            tmem_ld16(tcol + j_c * 16, old);
                            tmem_wait_ld();
                            if (sub != (kp_c % SUB)) {
#pragma unroll
                                for (int i = 0; i < 16; i++) nv[i] = old[i];
                            }
                        }
                        tmem_st16(tcol + j_c * 16, nv);
                        tmem_wait_st();
                    }
                    uint32_t r[64];
                    tmem_ld32(tcol + half * 64, r);
                    tmem_ld32(tcol + half * 64 + 32, r + 32);
                    const float4 q0 = *reinterpret_cast<const float4*>(qrow), q1 = *reinterpret_cast<const float4*>(qrow + 4);
                    const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                    tmem_wait_ld();
                    float sc[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float a = 0.f;
#pragma unroll
                        for (int d = 0; d < 8; d++) a = fmaf(qv[d], __uint_as_float(r[j * 16 + d]), a);
                        sc[j] = (full || sbase + j <= s) ? a : -INFINITY;
                    }
                    const float m = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
                    const float mu = (m == -INFINITY) ? 0.f : m;
                    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    float lsum = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float pj = ex2_approx(sc[j] - mu);
                        lsum += pj;
#pragma unroll
                        for (int d = 0; d < 8; d++) o[d] = fmaf(pj, __uint_as_float(r[j * 16 + 8 + d]), o[d]);
                    }
                    float* pp = sP + warp * 320 + lane;
                    pp[0] = m;
                    pp[32] = lsum;
#pragma unroll
                    for (int d = 0; d < 8; d++) pp[64 + 32 * d] = o[d];
                }
                TMEM_TICK(2)
                __syncthreads();
                // ---- merge the partials of each pair: thread = (channel, pair) ----
                {
                    const int ch = tid >> 5, pr = tid & 31;
                    if (pr < PAIRS) {
                        float mm[8 * SUB], ll[8 * SUB], oo[8 * SUB];
#pragma unroll
                        for (int i = 0; i < 8 * SUB; i++) {
                            const float* pp = sP + (i / SUB) * 320 + (i % SUB) * PAIRS + pr;
                            mm[i] = pp[0];
                            ll[i] = pp[32];
                            oo[i] = pp[64 + 32 * ch];
                        }
                        float M = mm[0];
#pragma unroll
                        for (int i = 1; i < 8 * SUB; i++) M = fmaxf(M, mm[i]);
                        float Ls = 0.f, O = 0.f;
#pragma unroll
                        for (int i = 0; i < 8 * SUB; i++) {
                            const float e = ex2_approx(mm[i] - M);
                            Ls = fmaf(ll[i], e, Ls);
                            O = fmaf(oo[i], e, O);
                        }
                        sAO[(pr >> 2) * LDH + (pr & 3) * 8 + ch] = __fdividef(O, Ls);
                    }
                }
                }
                TMEM_TICK(3)
                __syncthreads();
                // ---- phase C: output projection + residual on the FP32 pipe (8 x 1,024 MACs: one output per thread; an
                //      MMA version pays two dependent tensor-core latencies for 4 instructions) ----
                {
                    const float* ar = CACHE ? sAO + warp * LDH : sQKV + warp * LDQ + 64;   // no window: attention output = v
                    const float* wp = w + L.wproj;
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int k4 = 0; k4 < 8; k4++) {
                        float4 a4 = *reinterpret_cast<const float4*>(ar + 4 * k4);
                        if constexpr (!CACHE) {                        // the two K halves of the v row
                            const float4 u4 = *reinterpret_cast<const float4*>(ar + 32 + 4 * k4);
                            a4.x += u4.x; a4.y += u4.y; a4.z += u4.z; a4.w += u4.w;
                        }
                        const float av4[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int k = 4 * k4 + i;
                            acc[i] = fmaf(av4[i], wp[k * 32 + (lane ^ (((k >> 1) & 3) << 3))], acc[i]);
                        }
                    }
                    const float hm = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + w[L.bproj + lane] + sH[warp * LDH + lane];
                    sHm[warp * LDH + lane] = hm;
                    sA2[warp * LDH + lane] = ln_row(hm, w + L.ln2w, w + L.ln2b, p.eps, lane);     // the warp owns the row
                }
                TMEM_TICK(4)
                __syncthreads();
                // ---- phase D: LN2, this warp's 16 hidden units of the MLP and their share of the MLP-out product (K split
                //      over the warps) ----
                {
                    float hm[8];
                    load_row_frag(sHm, LDH, g, q, hm);
                    AFrag a[4];
                    frags_from_rows(sA2, LDH, g, q, a);
                    float fh[2][2][4], fl[2][2][4];                   // [tile][k pair]: chains of two
#pragma unroll
                    for (int k = 0; k < 4; k++)
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            float w0, w1;
                            ldb(w + L.wfc, 128, k, 2 * warp + j, g, q, w0, w1);
                            if (k < 2) mma2z(fh[j][k], fl[j][k], a[k], w0, w1);
                            else mma2(fh[j][k - 2], fl[j][k - 2], a[k], w0, w1);
                        }
                    AFrag fa[2];
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const float2 b2 = *reinterpret_cast<const float2*>(w + L.bfc + 8 * (2 * warp + j) + 2 * q);
                        float2 f = fold2(fh[j][0], fl[j][0], fh[j][1], fl[j][1]);
                        f.x += b2.x; f.y += b2.y;
                        gelu_new_sig2(f.x, f.y);
                        fa[j] = make_a(f.x, f.y);
                    }
                    float ph[4][4], pl[4][4];
#pragma unroll
                    for (int j = 0; j < 2; j++)
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            float w0, w1;
                            ldb(w + L.wpr, 32, 2 * warp + j, t, g, q, w0, w1);       // MLP-in n-tile = MLP-out k-step
                            if (j == 0) mma2z(ph[t], pl[t], fa[j], w0, w1);
                            else mma2(ph[t], pl[t], fa[j], w0, w1);
                        }
                    float* rr = sP + warp * (8 * LDR) + g * LDR + 2 * q;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        float2 y = fold(ph[t], pl[t]);
                        if (warp == 0) {
                            const float2 b2 = *reinterpret_cast<const float2*>(w + L.bpr + 8 * t + 2 * q);
                            y.x += hm[2 * t] + b2.x;
                            y.y += hm[2 * t + 1] + b2.y;
                        }
                        *reinterpret_cast<float2*>(rr + 8 * t) = y;
                    }
                }
                TMEM_TICK(5)
                __syncthreads();
                // ---- reduce the 8 partial rows: new residual ----
                {
                    const int r = tid >> 5, c = tid & 31;
                    float acc = 0.f;
#pragma unroll
                    for (int wv = 0; wv < 8; wv++) acc += sP[wv * (8 * LDR) + r * LDR + c];
                    sH[r * LDH + c] = acc;
                    // what reads the new residual next: LayerNorm 1 of the next layer, or the final LayerNorm of the head
                    const float* lw = (l + 1 < NL) ? w + L.layer_stride + L.ln1w : gw;
                    const float* lb = (l + 1 < NL) ? w + L.layer_stride + L.ln1b : gw + 32;
                    sA1[r * LDH + c] = ln_row(acc, lw, lb, p.eps, c);
                }
                __syncthreads();
                TMEM_TICK(6)
            }
            // ---- final LayerNorm + linear head: warp t < 4 computes output columns 8t .. 8t+8 ----
            if (warp < 4) {
                AFrag a[4];
                frags_from_rows(sA1, LDH, g, q, a);                     // final LayerNorm applied by the reduce phase
                float dh[4][4], dl[4][4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    float w0, w1;
                    ldb(gw + 64, 32, k, warp, g, q, w0, w1);
                    mma2z(dh[k], dl[k], a[k], w0, w1);
                }
                const int col = 8 * warp + 2 * q;
                const float2 b2 = *reinterpret_cast<const float2*>(gw + 64 + 1024 + col);
                const float2 ya = fold2(dh[0], dl[0], dh[1], dl[1]), yb = fold2(dh[2], dl[2], dh[3], dl[3]);
                const float y0 = ya.x + yb.x + b2.x, y1 = ya.y + yb.y + b2.y;
                const int b = grp * TRAJ + g;
                if (g < TRAJ && b < p.B)
                    *reinterpret_cast<float2*>(p.out + (size_t)b * p.out_stride + (size_t)s * 32 + col) = make_float2(y0, y1);
                const int pnext = CACHE ? min(s + 1, NCTX - 1) : 0;
                const float2 pe = __ldg(reinterpret_cast<const float2*>(p.pe + pnext * 32 + col));
                *reinterpret_cast<float2*>(sHn + g * LDH + col) = make_float2(y0 + pe.x, y1 + pe.y);
            }
            __syncthreads();
            { float* t = sH; sH = sHn; sHn = t; }
            sA1[warp * LDH + lane] = ln_row(sH[warp * LDH + lane], wsm + L.ln1w, wsm + L.ln1b, p.eps, lane);   // layer 0 of the next step
            __syncthreads();
            TMEM_TICK(7)
        }
        // ---- optional: the final windows in chronological order, [n_layer][2][B][H][Tk][hd] (attention.py:188) ----
        if (CACHE && p.presents != nullptr) {
            const int Tk = min(p.S, NCTX);
            const int b = grp * TRAJ + atraj;
            const size_t per_l = (size_t)2 * p.B * 4 * Tk * 8;
            for (int l = 0; l < NL; l++) {
                uint32_t r[64];
                tmem_ld32(tlane + l * 128 + half * 64, r);
                tmem_ld32(tlane + l * 128 + half * 64 + 32, r + 32);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int sidx = sbase + j;
                    const int t = (p.S >= NCTX) ? ((sidx - p.S) & (NCTX - 1)) : sidx;
                    if (b < p.B && t < Tk) {
#pragma unroll
                        for (int which = 0; which < 2; which++) {
                            float* dst = p.presents + (size_t)l * per_l + ((((size_t)which * p.B + b) * 4 + ahead) * Tk + t) * 8;
                            *reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(r[j * 16 + which * 8]), __uint_as_float(r[j * 16 + which * 8 + 1]),
                                                                          __uint_as_float(r[j * 16 + which * 8 + 2]), __uint_as_float(r[j * 16 + which * 8 + 3]));
                            *reinterpret_cast<float4*>(dst + 4) = make_float4(__uint_as_float(r[j * 16 + which * 8 + 4]), __uint_as_float(r[j * 16 + which * 8 + 5]),
                                                                              __uint_as_float(r[j * 16 + which * 8 + 6]), __uint_as_float(r[j * 16 + which * 8 + 7]));
                        }
                    }
                }
            }
        }
    }
    if constexpr (PROF) {
        if (tid == 0 && blockIdx.x == 0)
            for (int i = 0; i < 8; i++) prof[i] = pacc[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc512(tmem_slot);
}

// LDTM throughput probe: every warp streams its 4 layers x 64 columns `iters` times
__global__ void __launch_bounds__(TT, 1) tmem_ld_probe_kernel(int iters, unsigned* sink, long long* cycles) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc512(&tmem_slot);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tlane = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; i++) z[i] = threadIdx.x;
    for (int l = 0; l < 4; l++)
        for (int c = 0; c < 4; c++) tmem_st16(tlane + l * 128 + c * 16, z);
    tmem_wait_st();
    __syncthreads();
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        uint32_t r[64];
        tmem_ld32(tlane + (it & 3) * 128, r);
        tmem_ld32(tlane + (it & 3) * 128 + 32, r + 32);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 64; i++) acc ^= r[i];
    }
    const long long t1 = clock64();
    if (acc == 0x12345u) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc512(tmem_slot);
}

}  // namespace

int tmem_rollout_group(const Gpt2Layout& L) {
    if (L.E != 32 || L.n_head != 4 || L.n_layer > 4) return 0;
    if (L.n_ctx == 32) return 8;
    if (L.n_ctx == 64) return 4;
    return 0;
}

size_t tmem_rollout_smem(const Gpt2Layout& L) { return ((size_t)L.total + SCR_FLOATS) * sizeof(float); }

template <int NCTX, bool PROF, bool CACHE = true>
int launch_one(const TmemRolloutParams& p, int grid, size_t smem, long long* prof, cudaStream_t st) {
    TRPX_CUDA(cudaFuncSetAttribute(rollout_e32t_kernel<NCTX, PROF, CACHE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rollout_e32t_kernel<NCTX, PROF, CACHE><<<grid, TT, smem, st>>>(p, prof);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

int launch_rollout_tmem(const TmemRolloutParams& p, int grid, cudaStream_t st) {
    const size_t smem = tmem_rollout_smem(p.lay);
    if (!p.use_cache)
        return p.lay.n_ctx == 32 ? launch_one<32, false, false>(p, grid, smem, nullptr, st)
                                 : launch_one<64, false, false>(p, grid, smem, nullptr, st);
    const char* env = getenv("TRPX_TMEM_PROF");
    if (env != nullptr && env[0] == '1') {
        long long* prof = nullptr;
        TRPX_CUDA(cudaMalloc(&prof, 8 * sizeof(long long)));
        int rc = p.lay.n_ctx == 32 ? launch_one<32, true>(p, grid, smem, prof, st) : launch_one<64, true>(p, grid, smem, prof, st);
        if (rc) return rc;
        long long h[8];
        TRPX_CUDA(cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(prof);
        const int G = tmem_rollout_group(p.lay);
        const double n = (double)((((p.B + G - 1) / G) + grid - 1) / grid) * p.S * p.lay.n_layer;
        static const char* names[8] = {"A: LN1+QKV", "wait", "B: append+attention", "merge", "C: proj", "D: LN2+MLP", "reduce", "head(/layer)"};
        for (int i = 0; i < 8; i++) fprintf(stderr, "[tmem prof] %-22s %8.1f cycles per layer\n", names[i], (double)h[i] / n);
        return TRPX_OK;
    }
    return p.lay.n_ctx == 32 ? launch_one<32, false>(p, grid, smem, nullptr, st) : launch_one<64, false>(p, grid, smem, nullptr, st);
}

}  // namespace trpx

// tcgen05.ld bytes per clock per SM with 8 warps (what bounds the window read of K1t)
extern "C" int trpx_probe_tmem_ld(int device, float* bytes_per_clk_out) {
    using namespace trpx;
    TRPX_REQUIRE(bytes_per_clk_out != nullptr, "null output");
    int n = 0;
    TRPX_CUDA(cudaGetDeviceCount(&n));
    TRPX_REQUIRE(device >= 0 && device < n, "device %d out of range", device);
    DeviceGuard guard(device);
    unsigned* sink = nullptr;
    long long* cyc = nullptr;
    TRPX_CUDA(cudaMalloc(&sink, 16));
    TRPX_CUDA(cudaMalloc(&cyc, 16));
    const int iters = 4096;
    tmem_ld_probe_kernel<<<1, TT>>>(64, sink, cyc);
    tmem_ld_probe_kernel<<<1, TT>>>(iters, sink, cyc);
    long long c = 0;
    TRPX_CUDA(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
    cudaFree(sink);
    cudaFree(cyc);
    TRPX_CUDA(cudaGetLastError());
    *bytes_per_clk_out = c > 0 ? (float)((double)iters * TT * 64 * 4 / (double)c) : 0.f;
    return TRPX_OK;
}
