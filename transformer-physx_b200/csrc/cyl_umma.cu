// Cylinder decoder layers as tcgen05 (UMMA) implicit GEMMs on sm_100a.
// Reference: trphysx/embedding/embedding_cylinder.py:70-88 (recoveryNet: [Upsample x2 bilinear align_corners=True ->
// Conv2d 3x3 replicate -> ReLU] x4 -> Conv2d 3x3), :156-170 (recover: un-normalise, cylinder mask).
//
// Formulation.  A layer is  out[pix][co] = sum_{dy,dx,ci} P[pix + (dy-1)*PW + (dx-1)][ci] * w[co][ci][dy][dx]  where P is
// the (upsampled) replicate-padded layer input of a frame written as ONE flat sequence of padded pixels (PW = W + 2
// padded pixels per row, H + 2 rows per frame).  In that flat space a 3x3 tap is a shift of the
// pixel index, so the patch is staged ONCE in shared memory in the K-major no-swizzle core-matrix order
// [plane hi|lo][ci/4][pixel][4 floats] and every tap reuses it through the START ADDRESS of the UMMA shared-memory
// descriptor (scripts/probes/umma_tf32_probe.cu verifies row-shifted descriptors bit-exactly).  The three dx taps are
// folded into the N dimension:  D[g][dx][co] = sum_{dy,ci} P[g + (dy-1)*PW][ci] * w[co][ci][dy][dx]  is one MMA of
// N = 3*Cout per (dy, 8 input channels), and  out[f] = D[f-1][0] + D[f][1] + D[f+1][2]  is a +-1 pixel shift done in
// the epilogue with warp shuffles.  This cuts the A-operand traffic 3x: M128 x N x K8 TF32 MMAs are bound by the
// 128 B/cycle shared-memory operand fetch for N < 128 (measured: 45.7 / 56.1 / 96.1 cycles at N = 48 / 96 / 192).
//
// Precision.  FP32 parity (rel-L2 <= 1e-5) needs the error-compensated product a*w = a_hi*w_hi + a_hi*w_lo + a_lo*w_hi
// (TF32 hi / lo planes of both operands).  The tensor core truncates its FP32 accumulator once per instruction
// (probe: -2.4e-6 relative bias after 144 chained MMAs), so the dominant a_hi*w_hi chain gets its own TMEM columns
// (at most 48 instructions long) and the two small terms accumulate in a second column set; the epilogue adds them.
// Where 6*Cout <= 256 the a_hi pass computes both column sets in ONE MMA against the concatenated [w_hi | w_lo] rows.
//
// Roles (640 threads, one persistent CTA per SM): warps 0-7 epilogue (two groups of four; TMEM -> registers -> global),
// warp 8 issues the MMAs, warp 9 is the loader: one lane streams, per stage, the layer-input rows the stage needs and the stage's slice of
// the weight image into shared memory with TMA bulk copies (cp.async.bulk + mbarrier complete_tx), warps 10-19 build the
// patch from the staged rows (bilinear upsample + TF32 split, generic-proxy stores + fence.proxy.async).  mbarrier
// rings: a DEEP load ring (rows + weights of a K chunk are small; the loader runs 3-6 chunks ahead, which is what hides
// the ~1.5 us TMA round trip) with ld_full (TMA -> producers, MMA) / ld_empty (producers + tcgen05.commit -> loader), a
// two-slot patch ring with p_full (producers -> MMA) / p_empty (tcgen05.commit -> producers), and per 128-row block
// tfull (tcgen05.commit -> epilogue) / tempty (epilogue -> MMA).
#include "common.cuh"
#include "cyl_umma.cuh"
#include <algorithm>
#include <cstdlib>

using namespace trpx;

namespace {

constexpr int UM_EPI_WARPS = 8, UM_PROD_WARPS = 10;                       // two epilogue groups of 4 warps (one per TMEM lane quarter)
constexpr int UM_MMA_WARP = UM_EPI_WARPS, UM_LOAD_WARP = UM_EPI_WARPS + 1;
constexpr int UM_THREADS = (UM_EPI_WARPS + 2 + UM_PROD_WARPS) * 32;       // 640
constexpr int UM_NPT = UM_PROD_WARPS * 32;                                // producer threads
constexpr int UM_BLK = 126;                                               // useful rows of a 128-row M block (1 overlap each side)
constexpr int UM_CHUNK = 16;                                              // output channels per epilogue pass (4 for the 3-channel last layer)

// EPI_: 0 ReLU -> channels-last-4 (decoder, last strided encoder layer), 1 last decoder layer (un-normalise, mask, planar),
//       2 ReLU -> the NEXT strided layer's space-to-depth image, 3 plain planar output (last encoder layer)
// DY0_: first kernel row with non-zero weights (1 for the space-to-depth form of a stride-2 layer); TRIM_: the last row and
// column of the H x W grid are not outputs (space-to-depth images carry one extra block row / column); KS_: accumulator
// sets per block, K chunks dealt round-robin (the tensor core truncates its accumulator once per instruction, so a chain
// of n MMAs carries a ~n * 3e-8 relative bias: splitting long chains keeps the deep, narrow encoder layers at 1e-6)
template <int CIN_, int COUT_, int COUT_REAL_, int H_, int W_, bool UP_, int MB_, int NBUF_, int PS_, int LS_, int EPI_, int DY0_ = 0, int TRIM_ = 0, int KS_ = 1>
struct UmmaLayer {
    static constexpr int CIN = CIN_, COUT = COUT_, COUT_REAL = COUT_REAL_, H = H_, W = W_, MB = MB_;
    static constexpr int LS = LS_;                                        // load ring depth (layer-input rows + weight slice of a K chunk)
    static constexpr int PS = PS_;                                        // patch ring depth
    static constexpr int NBUF = NBUF_;                                    // TMEM accumulator buffers (2: the epilogue of item i overlaps the MMAs of item i+1)
    static constexpr int EPI = EPI_, DY0 = DY0_, TRIM = TRIM_, KS = KS_;
    static constexpr bool NARROW = (EPI_ == 1 || EPI_ == 3);              // <= 4 real output channels
    static constexpr int EC = NARROW ? 4 : UM_CHUNK;                      // output channels per epilogue pass
    static constexpr int NCHUNK = NARROW ? 1 : COUT_ / UM_CHUNK;          // epilogue passes per block
    static constexpr int EPI_GROUPS_PER_BLOCK = (NCHUNK % 2 == 0) ? 2 : 1;   // (block, chunk) units are dealt round-robin to the 2 groups
    static constexpr bool UP = UP_, FINAL = (EPI_ == 1);
    static constexpr int HS = UP ? H / 2 : H, WS = UP ? W / 2 : W;        // layer input size
    static constexpr int PW = W + 2, FRAME_POS = (H + 2) * PW;            // flat padded pixels per row / frame
    static constexpr int NPOS = UM_BLK * MB + 2 + 2 * PW;                 // staged positions of one work item
    static constexpr int ITEM_OUT = UM_BLK * MB;                          // output positions of one work item
    static constexpr int IPF = (H * PW + W + ITEM_OUT - 1) / ITEM_OUT;    // work items per frame (last output: row H, column W)
    static constexpr int ROWS_OUT = (NPOS + PW - 2) / PW + 1;             // padded rows a work item can touch
    static constexpr int SRC_ROWS = UP ? ROWS_OUT / 2 + 3 : ROWS_OUT;     // layer-input rows behind them
    static constexpr int SRC_BYTES = 2 * SRC_ROWS * WS * 16;              // [quad 2][row][col][4 f]
    static constexpr int N3 = 3 * COUT, N6 = 6 * COUT;
    static constexpr bool CONCAT = N6 <= 256;                             // a_hi . [w_hi | w_lo] in one MMA
    static constexpr int KCH = CIN / 8;                                   // K chunks of 8 input channels
    static constexpr int PATCH_BYTES = 2 * 2 * NPOS * 16;                 // [plane][kc][pos][4 f]
    static constexpr int W_BYTES = 3 * 2 * N6 * 16;                       // [dy][kc][n][4 f]
    static constexpr int LOAD_BYTES = W_BYTES + SRC_BYTES;                // one load-ring slot: [weights][rows]
    static constexpr int TMEM_USED = NBUF * MB * KS * N6;
    static constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    static constexpr int EDGE_FLOATS = 2 * 2 * 2 * 4 * UM_CHUNK;          // [group][parity][left|right][quarter][16]
    static constexpr size_t RING_BYTES = (size_t)PS * PATCH_BYTES + (size_t)LS * LOAD_BYTES;
    static constexpr size_t SMEM = RING_BYTES + EDGE_FLOATS * 4 + 512 + 128;
    static_assert(CIN % 8 == 0 && COUT % 16 == 0 && N3 % 16 == 0 && N3 <= 256, "shape");
    static_assert(TMEM_USED <= 512, "TMEM");
    static_assert(UM_BLK * (MB - 1) + 2 * PW + 128 <= NPOS, "patch too small");
};

struct UmmaArgs {
    const float* in;        // [N][CIN/4][HS][WS][4]
    const float* wimg;      // [KCH][dy 3][kc 2][N6][4]
    const float* bias;      // [COUT_REAL]
    float* out;
    const float* mu;
    const float* sd;
    const unsigned char* mask;
    int N;
    int nitems;
    int co_base, co_total;  // EPI 0: the COUT channels of this launch are channels [co_base, co_base + COUT) of co_total
    int debug;              // TRPX_UMMA_DEBUG bits (timing experiments only, results are garbage): 1 producers idle, 2 no MMAs,
                            // 4 epilogue skips the stores, 8 no fence.proxy.async, 16 no layer-input TMA, 32 no weight TMA,
                            // 64 epilogue skips TMEM loads and the edge exchange
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
    // K-major, no swizzle: start >> 4 | LBO >> 4 (stride between the two 16-byte K chunks) | SBO = 128 B (stride between
    // 8-row groups) | descriptor version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    // D = F32 (1 @4), A = B = TF32 (2 @7, 2 @10), both K-major, N >> 3 @17, M >> 4 @24
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    hi = __uint_as_float(u);
    lo = x - hi;
}

// packed FP32 (f32x2) helpers: (d0, d1) = (a0, a1) * b   and   (d0, d1) = (a0, a1) - (b0, b1)
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b) {
    asm("{\n.reg .b64 ra, rb, rc;\n"
        "mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %4};\n"
        "mul.rn.f32x2 rc, ra, rb;\n"
        "mov.b64 {%0, %1}, rc;\n}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b));
}
__device__ __forceinline__ void fsub2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n.reg .b64 ra, rb, rc;\n"
        "mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n"
        "sub.rn.f32x2 rc, ra, rb;\n"
        "mov.b64 {%0, %1}, rc;\n}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// bilinear x2 (align_corners=True) source index / weight, as aten's area_pixel_compute_source_index
struct Lerp1 {
    int i0, i1;
    float l1;
};
__device__ __forceinline__ Lerp1 lerp1(int dst, int in_size, float scale) {
    const float real = scale * (float)dst;
    Lerp1 r;
    r.i0 = min((int)real, in_size - 1);
    r.l1 = fminf(fmaxf(real - (float)r.i0, 0.f), 1.f);
    r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
    return r;
}

// layer-input rows [i_lo, i_lo + nrows) needed by work item j of a frame (the loader and the producers agree on this)
template <class L>
__device__ __forceinline__ void item_src_rows(int lb, int& i_lo, int& nrows) {
    const int lp0 = max(lb - L::PW, 0), lp1 = min(lb - L::PW + L::NPOS - 1, L::FRAME_POS - 1);
    const int y0 = min(max(lp0 / L::PW - 1, 0), L::H - 1), y1 = min(max(lp1 / L::PW - 1, 0), L::H - 1);
    if constexpr (L::UP) {
        const float sh = (float)(L::HS - 1) / (float)(L::H - 1);
        i_lo = lerp1(y0, L::HS, sh).i0;
        nrows = lerp1(y1, L::HS, sh).i1 - i_lo + 1;
    } else {
        i_lo = y0;
        nrows = y1 - y0 + 1;
    }
    nrows = min(nrows, L::SRC_ROWS);
}

template <class L>
__global__ void __launch_bounds__(UM_THREADS, 1) cyl_umma_conv_kernel(const UmmaArgs a) {
    extern __shared__ __align__(128) unsigned char um_smem[];
    unsigned char* patches = um_smem;                                    // [PS][PATCH_BYTES]
    unsigned char* loads = um_smem + (size_t)L::PS * L::PATCH_BYTES;     // [LS][weights | layer-input rows]
    float* edge = reinterpret_cast<float*>(um_smem + L::RING_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(edge + L::EDGE_FLOATS);
    uint64_t* p_full = bars;                     // [PS] patch built (producer warps)
    uint64_t* p_empty = p_full + L::PS;          // [PS] patch consumed (tcgen05.commit)
    uint64_t* ld_full = p_empty + L::PS;         // [LS] rows + weights landed (TMA complete_tx)
    uint64_t* ld_empty = ld_full + L::LS;        // [LS] rows read by the producers + weights consumed (tcgen05.commit)
    uint64_t* tfull = ld_empty + L::LS;          // [NBUF][MB] accumulator of block b complete
    uint64_t* tempty = tfull + L::NBUF * L::MB;  // [NBUF][MB] accumulator of block b drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + L::NBUF * L::MB);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < L::PS; s++) {
            mbar_init(&p_full[s], UM_PROD_WARPS);
            mbar_init(&p_empty[s], 1);
        }
        for (int s = 0; s < L::LS; s++) {
            mbar_init(&ld_full[s], 1);
            mbar_init(&ld_empty[s], UM_PROD_WARPS + 1);
        }
        for (int b = 0; b < L::NBUF * L::MB; b++) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4 * L::EPI_GROUPS_PER_BLOCK);
        }
    }
    if (warp == UM_MMA_WARP) tmem_alloc(tmem_slot, L::TMEM_COLS);
    // staged positions outside the frame (first / last work item of a frame) are never written: start from zeros so
    // that the rows of the accumulator nobody reads are at least finite
    {
        float4* pz = reinterpret_cast<float4*>(patches);
        for (int i = tid; i < L::PS * L::PATCH_BYTES / 16; i += UM_THREADS) pz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < UM_EPI_WARPS) {
        // =============================== epilogue ===============================
        // two groups of four warps (warp % 4 = TMEM lane quarter); the (block, 16-channel chunk) units of an item are
        // dealt round-robin to the groups
        const int eg = warp >> 2, q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int par = 0;
        int iter = 0;
        for (int item = blockIdx.x; item < a.nitems; item += gridDim.x, iter++) {
            const int n = item / L::IPF;
            const int lb = (item - n * L::IPF) * L::ITEM_OUT;
            const int buf = L::NBUF == 2 ? (iter & 1) : 0;
            const uint32_t tpar = L::NBUF == 2 ? ((iter >> 1) & 1) : (iter & 1);
            for (int b = 0; b < L::MB; b++) {
                if (L::EPI_GROUPS_PER_BLOCK == 1 && ((b * L::NCHUNK) & 1) != eg) continue;       // NCHUNK odd: whole block to one group
                mbar_wait_backoff(&tfull[buf * L::MB + b], tpar, 32);
                tc_fence_after();
                const int t = q * 32 + lane;                           // row of the block
                const int lp = lb + UM_BLK * b + t;                    // flat padded position inside the frame
                const int yp = lp / L::PW, xp = lp - yp * L::PW;
                const bool valid = t >= 1 && t <= UM_BLK && yp >= 1 && yp <= L::H - L::TRIM && xp >= 1 && xp <= L::W - L::TRIM;
                const int y = yp - 1, x = xp - 1;
                const uint32_t tcol = tmem_base + lane_base + (uint32_t)((buf * L::MB + b) * L::KS * L::N6);
                if (a.debug & 64) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[buf * L::MB + b]);
                    continue;
                }
                constexpr int CSTEP = L::EPI_GROUPS_PER_BLOCK;        // 2: this group takes every other chunk of the block
                constexpr int EC = L::EC;
                for (int c = (CSTEP == 2 ? eg : 0); c < L::NCHUNK; c += CSTEP) {
                    const int co0 = c * EC;
                    float v[3][EC];
#pragma unroll
                    for (int dx = 0; dx < 3; dx++) {
#pragma unroll
                        for (int j = 0; j < EC; j++) v[dx][j] = 0.f;
#pragma unroll
                        for (int ks = 0; ks < L::KS; ks++) {
                            uint32_t rh[EC], rs[EC];
                            if constexpr (EC == 16) {
                                tmem_ld16(tcol + (uint32_t)(ks * L::N6 + dx * L::COUT + co0), rh);
                                tmem_ld16(tcol + (uint32_t)(ks * L::N6 + L::N3 + dx * L::COUT + co0), rs);
                            } else {
                                tmem_ld4(tcol + (uint32_t)(ks * L::N6 + dx * L::COUT + co0), rh);
                                tmem_ld4(tcol + (uint32_t)(ks * L::N6 + L::N3 + dx * L::COUT + co0), rs);
                            }
                            tmem_wait_ld();
#pragma unroll
                            for (int j = 0; j < EC; j++) v[dx][j] += __uint_as_float(rh[j]) + __uint_as_float(rs[j]);
                        }
                    }
                    if (c + CSTEP >= L::NCHUNK) {                      // this group is done with the block's accumulator
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[buf * L::MB + b]);
                    }
                    // rows 31|32, 63|64, 95|96 of the block sit in different warps: their dx = 0 / dx = 2 partial sums travel
                    // through shared memory (one float4 vector store / load per 4 channels)
                    float4* eb = reinterpret_cast<float4*>(edge + (size_t)((eg * 2 + par) * 2) * 4 * UM_CHUNK);
                    if (lane == 31) {
#pragma unroll
                        for (int j = 0; j < EC / 4; j++) eb[(0 * 4 + q) * (UM_CHUNK / 4) + j] = make_float4(v[0][4 * j], v[0][4 * j + 1], v[0][4 * j + 2], v[0][4 * j + 3]);
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < EC / 4; j++) eb[(1 * 4 + q) * (UM_CHUNK / 4) + j] = make_float4(v[2][4 * j], v[2][4 * j + 1], v[2][4 * j + 2], v[2][4 * j + 3]);
                    }
                    named_bar_sync(1 + eg, 128);
                    float lv[EC], rv[EC], o[EC];
#pragma unroll
                    for (int j = 0; j < EC; j++) {
                        lv[j] = __shfl_up_sync(FULL, v[0][j], 1);
                        rv[j] = __shfl_down_sync(FULL, v[2][j], 1);
                    }
                    if (lane == 0 && q > 0) {
#pragma unroll
                        for (int j = 0; j < EC / 4; j++) {
                            const float4 e = eb[(0 * 4 + q - 1) * (UM_CHUNK / 4) + j];
                            lv[4 * j] = e.x; lv[4 * j + 1] = e.y; lv[4 * j + 2] = e.z; lv[4 * j + 3] = e.w;
                        }
                    }
                    if (lane == 31 && q < 3) {
#pragma unroll
                        for (int j = 0; j < EC / 4; j++) {
                            const float4 e = eb[(1 * 4 + q + 1) * (UM_CHUNK / 4) + j];
                            rv[4 * j] = e.x; rv[4 * j + 1] = e.y; rv[4 * j + 2] = e.z; rv[4 * j + 3] = e.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < EC; j++) o[j] = (lv[j] + v[1][j]) + rv[j];
                    par ^= 1;
                    if (valid && !(a.debug & 4)) {
                        if constexpr (L::EPI == 0) {
                            float4* out4 = reinterpret_cast<float4*>(a.out);
#pragma unroll
                            for (int c4 = 0; c4 < EC / 4; c4++) {
                                const int co = co0 + 4 * c4;
                                const float4 bz = __ldg(reinterpret_cast<const float4*>(a.bias + co));
                                float4 w;
                                w.x = fmaxf(o[4 * c4 + 0] + bz.x, 0.f);
                                w.y = fmaxf(o[4 * c4 + 1] + bz.y, 0.f);
                                w.z = fmaxf(o[4 * c4 + 2] + bz.z, 0.f);
                                w.w = fmaxf(o[4 * c4 + 3] + bz.w, 0.f);
                                out4[(((size_t)n * (a.co_total / 4) + ((a.co_base + co) >> 2)) * (L::H - L::TRIM) + y) * (L::W - L::TRIM) + x] = w;
                            }
                        } else if constexpr (L::EPI == 2) {
                            // the next layer is a stride-2 conv: write its space-to-depth image Q[(py,px,co)][Y][X] =
                            // out[2Y+py-1][2X+px-1][co] (replicate: the border rows / columns are stored twice)
                            constexpr int HO = L::H - 1, WO = L::W - 1, H2 = HO / 2 + 1, W2 = WO / 2 + 1;
                            const int QP = a.co_total / 4;                     // quads per parity of the next layer's image
                            float4* out4 = reinterpret_cast<float4*>(a.out);
                            int Ys[2], pys[2], ny = 1, Xs[2], pxs[2], nx = 1;
                            Ys[0] = (y + 1) >> 1; pys[0] = (y + 1) & 1;
                            Xs[0] = (x + 1) >> 1; pxs[0] = (x + 1) & 1;
                            if (y == 0) { Ys[1] = 0; pys[1] = 0; ny = 2; }
                            if (y == HO - 1) { Ys[1] = HO / 2; pys[1] = 1; ny = 2; }
                            if (x == 0) { Xs[1] = 0; pxs[1] = 0; nx = 2; }
                            if (x == WO - 1) { Xs[1] = WO / 2; pxs[1] = 1; nx = 2; }
#pragma unroll
                            for (int c4 = 0; c4 < EC / 4; c4++) {
                                const int co = co0 + 4 * c4;
                                const float4 bz = __ldg(reinterpret_cast<const float4*>(a.bias + co));
                                float4 w;
                                w.x = fmaxf(o[4 * c4 + 0] + bz.x, 0.f);
                                w.y = fmaxf(o[4 * c4 + 1] + bz.y, 0.f);
                                w.z = fmaxf(o[4 * c4 + 2] + bz.z, 0.f);
                                w.w = fmaxf(o[4 * c4 + 3] + bz.w, 0.f);
                                for (int iy = 0; iy < ny; iy++)
                                    for (int ix = 0; ix < nx; ix++) {
                                        const int quad = (pys[iy] * 2 + pxs[ix]) * QP + ((a.co_base + co) >> 2);
                                        out4[(((size_t)n * (4 * QP) + quad) * H2 + Ys[iy]) * W2 + Xs[ix]] = w;
                                    }
                            }
                        } else if constexpr (L::EPI == 3) {
#pragma unroll
                            for (int j = 0; j < L::COUT_REAL; j++)
                                a.out[(((size_t)n * L::COUT_REAL + j) * L::H + y) * L::W + x] = o[j] + __ldg(a.bias + j);
                        } else {
                            const bool masked = a.mask != nullptr && a.mask[y * L::W + x] != 0;
#pragma unroll
                            for (int j = 0; j < L::COUT_REAL; j++) {
                                float w = o[j] + __ldg(a.bias + j);
                                w = __ldg(a.sd + j) * w + __ldg(a.mu + j);        // embedding_cylinder.py:202-203
                                if (masked) w = 0.f;                               // :168-169
                                a.out[(((size_t)n * L::COUT_REAL + j) * L::H + y) * L::W + x] = w;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == UM_MMA_WARP) {
        // =============================== MMA issue ===============================
        if (lane == 0) {
            constexpr uint32_t idesc6 = umma_idesc_tf32(128, L::CONCAT ? L::N6 : L::N3);
            constexpr uint32_t idesc3 = umma_idesc_tf32(128, L::N3);
            constexpr uint32_t LBO_A = L::NPOS * 16, LBO_B = L::N6 * 16;
            int ps = 0, pph = 0, ls = 0, lph = 0, iter = 0;
            for (int item = blockIdx.x; item < a.nitems; item += gridDim.x, iter++) {
                for (int kc = 0; kc < L::KCH; kc++) {
                    mbar_wait(&ld_full[ls], lph);
                    mbar_wait(&p_full[ps], pph);
                    tc_fence_after();
                    const uint32_t sp = smem_u32(patches + (size_t)ps * L::PATCH_BYTES);
                    const uint32_t sw = smem_u32(loads + (size_t)ls * L::LOAD_BYTES);
                    const int buf = L::NBUF == 2 ? (iter & 1) : 0;
                    const uint32_t tpar = L::NBUF == 2 ? ((iter >> 1) & 1) : (iter & 1);
                    for (int b = 0; b < L::MB; b++) {
                        if (kc == 0) {
                            mbar_wait(&tempty[buf * L::MB + b], tpar ^ 1);
                            tc_fence_after();
                        }
                        const uint32_t tc = tmem_base + (uint32_t)(((buf * L::MB + b) * L::KS + (kc % L::KS)) * L::N6);
#pragma unroll
                        for (int dy = L::DY0; dy < 3; dy++) {
                            if (a.debug & 2) break;
                            const uint32_t arow = (uint32_t)(UM_BLK * b + dy * L::PW) * 16u;
                            const uint64_t a_hi = umma_desc(sp + arow, LBO_A);
                            const uint64_t a_lo = umma_desc(sp + 2 * L::NPOS * 16 + arow, LBO_A);
                            const uint32_t wb = sw + (uint32_t)dy * (2 * L::N6 * 16);
                            const uint32_t first = (kc < L::KS && dy == L::DY0) ? 0u : 1u;
                            if constexpr (L::CONCAT) {
                                umma_tf32(tc, a_hi, umma_desc(wb, LBO_B), idesc6, first);                  // hh | hl
                                umma_tf32(tc + L::N3, a_lo, umma_desc(wb, LBO_B), idesc3, 1u);              // + lh
                            } else {
                                umma_tf32(tc, a_hi, umma_desc(wb, LBO_B), idesc3, first);                   // hh
                                umma_tf32(tc + L::N3, a_hi, umma_desc(wb + L::N3 * 16, LBO_B), idesc3, first);   // hl
                                umma_tf32(tc + L::N3, a_lo, umma_desc(wb, LBO_B), idesc3, 1u);              // + lh
                            }
                        }
                        if (kc == L::KCH - 1) umma_commit(&tfull[buf * L::MB + b]);
                    }
                    umma_commit(&p_empty[ps]);
                    umma_commit(&ld_empty[ls]);
                    if (++ps == L::PS) { ps = 0; pph ^= 1; }
                    if (++ls == L::LS) { ls = 0; lph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == UM_LOAD_WARP) {
        // =============================== loader (TMA bulk copies) ===============================
        if (lane == 0) {
            int s = 0, ph = 0;
            for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
                const int n = item / L::IPF;
                const int lb = (item - n * L::IPF) * L::ITEM_OUT;
                int i_lo, nrows;
                item_src_rows<L>(lb, i_lo, nrows);
                const uint32_t row_bytes = (uint32_t)nrows * L::WS * 16u;
                for (int kc = 0; kc < L::KCH; kc++) {
                    mbar_wait_backoff(&ld_empty[s], ph ^ 1, 40);
                    unsigned char* st = loads + (size_t)s * L::LOAD_BYTES;
                    const bool do_src = !(a.debug & 16), do_w = !(a.debug & 32);
                    mbar_expect_tx(&ld_full[s], (do_src ? 2 * row_bytes : 0) + (do_w ? L::W_BYTES : 0));
#pragma unroll
                    for (int qd = 0; qd < 2; qd++) {
                        const float* src = a.in + ((((size_t)n * (L::CIN / 4) + 2 * kc + qd) * L::HS + i_lo) * L::WS) * 4;
                        if (do_src) bulk_g2s(st + L::W_BYTES + qd * (L::SRC_ROWS * L::WS * 16), src, row_bytes, &ld_full[s]);
                    }
                    if (do_w) bulk_g2s(st, a.wimg + (size_t)kc * (L::W_BYTES / 4), L::W_BYTES, &ld_full[s]);
                    if (++s == L::LS) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // =============================== producers ===============================
        // One thread owns ONE padded column xp (both channel quads of the chunk: two independent dependency chains) and
        // a slice of RPT rows of the item.  Everything that depends only on the position (source rows, interpolation
        // weights, patch offset) is computed once per item; per chunk a row costs, per quad, 4 packed FMAs for the vertical
        // interpolation, 4 cvt + 2 packed subtractions for the TF32 split and 2 STS.128, plus 2 LDS.128 + 4 packed FMAs
        // whenever the source row advances (every other output row).  The producers are latency bound (few warps), so
        // instructions per thread and independent chains per thread are what count.
        const int ptid = tid - (UM_LOAD_WARP + 1) * 32;
        constexpr int RSPLIT_ = UM_NPT / L::PW;
        constexpr int RSPLIT = RSPLIT_ < 1 ? 1 : (RSPLIT_ > L::ROWS_OUT ? L::ROWS_OUT : RSPLIT_);
        constexpr int RPT = (L::ROWS_OUT + RSPLIT - 1) / RSPLIT;           // rows per thread
        static_assert(L::PW <= UM_NPT, "one producer thread per padded column");
        const bool active = ptid < L::PW * RSPLIT;
        const int xp = ptid % L::PW, rs = ptid / L::PW;
        const int xsrc = min(max(xp - 1, 0), L::W - 1);                    // replicate padding
        const float sh = L::UP ? (float)(L::HS - 1) / (float)(L::H - 1) : 0.f;
        int j0 = xsrc, dj = 0;
        float wx1 = 0.f;
        if constexpr (L::UP) {
            const Lerp1 lx = lerp1(xsrc, L::WS, (float)(L::WS - 1) / (float)(L::W - 1));
            j0 = lx.i0; dj = lx.i1 - lx.i0; wx1 = lx.l1;
        }
        const float wx0 = 1.f - wx1;
        constexpr int QS = L::SRC_ROWS * L::WS;                            // float4 per staged quad
        int ps = 0, pph = 0, ls = 0, lph = 0;
        for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
            const int n = item / L::IPF;
            const int lb = (item - n * L::IPF) * L::ITEM_OUT;
            int i_lo, nrows;
            item_src_rows<L>(lb, i_lo, nrows);
            const int lp_a = max(lb - L::PW, 0), lp_b = min(lb - L::PW + L::NPOS, L::FRAME_POS);     // staged flat range
            const int yp_a = lp_a / L::PW, yp_b = (lp_b - 1) / L::PW;
            // per-row records of this thread (rows yp_a + rs*RPT ...): float4 offsets of the two source rows in the staged
            // quad (rowA < 0: nothing to do), vertical weight, byte offset in the patch
            int rowA[RPT], rowB[RPT];
            uint32_t poff[RPT];
            float w1[RPT];
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                const int yp = yp_a + rs * RPT + r;
                const int p = yp * L::PW + xp - (lb - L::PW);
                rowA[r] = -1; rowB[r] = 0; w1[r] = 0.f; poff[r] = 0;
                if (active && yp <= yp_b && p >= 0 && p < L::NPOS) {
                    const int y = min(max(yp - 1, 0), L::H - 1);
                    if constexpr (L::UP) {
                        const Lerp1 ly = lerp1(y, L::HS, sh);
                        rowA[r] = (ly.i0 - i_lo) * L::WS + j0;
                        rowB[r] = (min(ly.i0 + 1, L::HS - 1) - i_lo) * L::WS + j0;
                        w1[r] = ly.l1;
                    } else {
                        rowA[r] = (y - i_lo) * L::WS + j0;
                    }
                    poff[r] = (uint32_t)p * 16u;
                }
            }
            for (int kc = 0; kc < L::KCH; kc++) {
                mbar_wait_backoff(&ld_full[ls], lph, 40);
                mbar_wait_backoff(&p_empty[ps], pph ^ 1, 40);
                if (!(a.debug & 1)) {
                    const float4* src4 = reinterpret_cast<const float4*>(loads + (size_t)ls * L::LOAD_BYTES + L::W_BYTES);
                    const uint32_t dbase = smem_u32(patches + (size_t)ps * L::PATCH_BYTES);
                    int cur = -1;
                    float hA[2][4], hB[2][4];
#pragma unroll
                    for (int r = 0; r < RPT; r++) {
                        if (rowA[r] < 0) continue;
                        float v[2][4];
                        if constexpr (L::UP) {
                            if (rowA[r] != cur) {
                                const bool shift = rowA[r] == cur + L::WS && cur >= 0;
#pragma unroll
                                for (int qd = 0; qd < 2; qd++) {
                                    if (shift) {
#pragma unroll
                                        for (int e = 0; e < 4; e++) hA[qd][e] = hB[qd][e];
                                    } else {
                                        const float4 u0 = src4[qd * QS + rowA[r]], u1 = src4[qd * QS + rowA[r] + dj];
                                        fmul2(hA[qd][0], hA[qd][1], u0.x, u0.y, wx0);
                                        fmul2(hA[qd][2], hA[qd][3], u0.z, u0.w, wx0);
                                        ffma2(hA[qd][0], hA[qd][1], u1.x, u1.y, wx1);
                                        ffma2(hA[qd][2], hA[qd][3], u1.z, u1.w, wx1);
                                    }
                                    const float4 t0 = src4[qd * QS + rowB[r]], t1 = src4[qd * QS + rowB[r] + dj];
                                    fmul2(hB[qd][0], hB[qd][1], t0.x, t0.y, wx0);
                                    fmul2(hB[qd][2], hB[qd][3], t0.z, t0.w, wx0);
                                    ffma2(hB[qd][0], hB[qd][1], t1.x, t1.y, wx1);
                                    ffma2(hB[qd][2], hB[qd][3], t1.z, t1.w, wx1);
                                }
                                cur = rowA[r];
                            }
                            const float wy1 = w1[r], wy0 = 1.f - wy1;
#pragma unroll
                            for (int qd = 0; qd < 2; qd++) {
                                fmul2(v[qd][0], v[qd][1], hA[qd][0], hA[qd][1], wy0);
                                fmul2(v[qd][2], v[qd][3], hA[qd][2], hA[qd][3], wy0);
                                ffma2(v[qd][0], v[qd][1], hB[qd][0], hB[qd][1], wy1);
                                ffma2(v[qd][2], v[qd][3], hB[qd][2], hB[qd][3], wy1);
                            }
                        } else {
#pragma unroll
                            for (int qd = 0; qd < 2; qd++) {
                                const float4 u = src4[qd * QS + rowA[r]];
                                v[qd][0] = u.x; v[qd][1] = u.y; v[qd][2] = u.z; v[qd][3] = u.w;
                            }
                        }
#pragma unroll
                        for (int qd = 0; qd < 2; qd++) {
                            float h[4], l[4];
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                uint32_t u;
                                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v[qd][e]));
                                h[e] = __uint_as_float(u);
                            }
                            fsub2(l[0], l[1], v[qd][0], v[qd][1], h[0], h[1]);
                            fsub2(l[2], l[3], v[qd][2], v[qd][3], h[2], h[3]);
                            const uint32_t d = dbase + (uint32_t)(qd * L::NPOS) * 16u + poff[r];
                            st_shared_v4(d, h[0], h[1], h[2], h[3]);
                            st_shared_v4(d + 2 * L::NPOS * 16, l[0], l[1], l[2], l[3]);
                        }
                    }
                }
                if (!(a.debug & 8)) fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&p_full[ps]);
                    mbar_arrive(&ld_empty[ls]);       // the staged rows have been read
                }
                if (++ps == L::PS) { ps = 0; pph ^= 1; }
                if (++ls == L::LS) { ls = 0; lph ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == UM_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, L::TMEM_COLS);
    }
}

// weight image: [kc8][dy][kc 2][n 6*COUT][4]; n = plane*3*COUT + dx*COUT + co; zero for co >= cout_real
__global__ void umma_wimg_kernel(const float* __restrict__ w, float* __restrict__ img, int CIN, int COUT, int cout_real) {
    const int N6 = 6 * COUT, N3 = 3 * COUT;
    const int total = (CIN / 8) * 3 * 2 * N6 * 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int r = i;
        const int e = r % 4; r /= 4;
        const int n = r % N6; r /= N6;
        const int kc = r % 2; r /= 2;
        const int dy = r % 3; const int k8 = r / 3;
        const int plane = n / N3, m = n % N3, dx = m / COUT, co = m % COUT;
        const int ci = k8 * 8 + kc * 4 + e;
        float v = 0.f;
        if (co < cout_real) v = w[((size_t)co * CIN + ci) * 9 + dy * 3 + dx];
        float hi, lo;
        split_tf32(v, hi, lo);
        img[i] = plane ? lo : hi;
    }
}

// Weight image of a stride-2 layer in its space-to-depth form: the layer reads Q[(py,px,ci)][Y][X] = in[2Y+py-1][2X+px-1][ci]
// with a 2x2 kernel (DY, DX in {0,1}), written here as the 3x3 image the kernel expects with a zero first row / column:
// tap (dy', dx') = (DY+1, DX+1) of channel k = (py*2+px)*Cin + ci carries w[co][ci][2DY+py][2DX+px] (zero beyond the 3x3).
__global__ void umma_wimg_s2d_kernel(const float* __restrict__ w, float* __restrict__ img, int Cin, int COUT, int co_base, int Cout_all) {
    const int CINQ = 4 * Cin, N6 = 6 * COUT, N3 = 3 * COUT;
    const int total = (CINQ / 8) * 3 * 2 * N6 * 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int r = i;
        const int e = r % 4; r /= 4;
        const int n = r % N6; r /= N6;
        const int kc = r % 2; r /= 2;
        const int dyp = r % 3; const int k8 = r / 3;
        const int plane = n / N3, m = n % N3, dxp = m / COUT, co = m % COUT;
        const int k = k8 * 8 + kc * 4 + e;
        const int par = k / Cin, ci = k % Cin, py = par >> 1, px = par & 1;
        float v = 0.f;
        if (dyp >= 1 && dxp >= 1) {
            const int ry = 2 * (dyp - 1) + py, rx = 2 * (dxp - 1) + px;
            if (ry <= 2 && rx <= 2 && co_base + co < Cout_all) v = w[((size_t)(co_base + co) * Cin + ci) * 9 + ry * 3 + rx];
        }
        float hi, lo;
        split_tf32(v, hi, lo);
        img[i] = plane ? lo : hi;
    }
}

// encoder input: cat(x, visc) -> (x - mu) / std (embedding_cylinder.py:149-150, 198-200), written as the space-to-depth
// image of the first stride-2 layer: Q[n][parity][33][65][4 channels]
__global__ void umma_enc_prep_kernel(const float* __restrict__ x, const float* __restrict__ visc, const float* __restrict__ mu,
                                     const float* __restrict__ sd, float4* __restrict__ q, int N) {
    constexpr int H = 64, W = 128, H2 = H / 2 + 1, W2 = W / 2 + 1;
    const long long total = (long long)N * 4 * H2 * W2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int X = (int)(i % W2), Y = (int)((i / W2) % H2), par = (int)((i / ((long long)W2 * H2)) % 4);
        const int n = (int)(i / (4LL * W2 * H2));
        const int yy = min(max(2 * Y + (par >> 1) - 1, 0), H - 1), xx = min(max(2 * X + (par & 1) - 1, 0), W - 1);
        const float* xp = x + (size_t)n * 3 * H * W + yy * W + xx;
        float4 v;
        v.x = (xp[0] - __ldg(mu + 0)) / __ldg(sd + 0);
        v.y = (xp[H * W] - __ldg(mu + 1)) / __ldg(sd + 1);
        v.z = (xp[2 * H * W] - __ldg(mu + 2)) / __ldg(sd + 2);
        v.w = (visc[n] - __ldg(mu + 3)) / __ldg(sd + 3);
        q[i] = v;
    }
}

//                      CIN COUT real  H    W   UP    MB NBUF PS LS EPI [DY0 TRIM]
using UL0 = UmmaLayer<128, 64, 64, 16, 32, true, 1, 1, 3, 4, 0>;
using UL1 = UmmaLayer<64, 32, 32, 32, 64, true, 2, 1, 3, 5, 0>;
using UL2 = UmmaLayer<32, 16, 16, 64, 128, true, 4, 1, 3, 3, 0>;
using UL3 = UmmaLayer<16, 16, 3, 64, 128, false, 4, 1, 2, 3, 1>;
// encoder (embedding_cylinder.py:42-60): four stride-2 layers in space-to-depth form (CIN = 4 x input channels, H x W =
// the (Ho+1) x (Wo+1) block grid), then the stride-1 128 -> 4 layer
using UE1 = UmmaLayer<16, 16, 16, 33, 65, false, 4, 1, 2, 3, 2, 1, 1>;      //   4 -> 16, 64x128 -> 32x64
using UE2 = UmmaLayer<64, 32, 32, 17, 33, false, 1, 1, 3, 4, 2, 1, 1, 2>;   //  16 -> 32, -> 16x32
using UE3 = UmmaLayer<128, 32, 32, 9, 17, false, 1, 1, 3, 4, 2, 1, 1, 2>;   //  32 -> 64 (two launches of 32 channels), -> 8x16
using UE4 = UmmaLayer<256, 32, 32, 5, 9, false, 1, 1, 3, 4, 0, 1, 1, 2>;    //  64 -> 128 (four launches of 32 channels), -> 4x8
using UE5 = UmmaLayer<128, 16, 4, 4, 8, false, 1, 1, 3, 4, 3, 0, 0, 4>;     // 128 -> 4 at 4x8, stride 1
static_assert(UE1::SMEM <= 227 * 1024 && UE2::SMEM <= 227 * 1024 && UE3::SMEM <= 227 * 1024 && UE4::SMEM <= 227 * 1024 &&
              UE5::SMEM <= 227 * 1024, "smem");
static_assert(UL0::SMEM <= 227 * 1024 && UL1::SMEM <= 227 * 1024 && UL2::SMEM <= 227 * 1024 && UL3::SMEM <= 227 * 1024, "smem");

template <class L>
int launch_umma(const UmmaArgs& base, int sm_count, cudaStream_t st) {
    UmmaArgs a = base;
    TRPX_REQUIRE((long long)a.N * L::IPF < (1LL << 30), "too many frames in one pass (%d)", a.N);
    {
        const char* dbg = getenv("TRPX_UMMA_DEBUG");
        a.debug = dbg ? atoi(dbg) : 0;
    }
    a.nitems = a.N * L::IPF;     // item j of a frame covers its flat output positions [j*126*MB + 1, (j+1)*126*MB]
    static_assert(L::SMEM <= 227 * 1024, "shared memory");
    TRPX_CUDA(cudaFuncSetAttribute(cyl_umma_conv_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
    const int grid = std::min(a.nitems, sm_count);
    if (getenv("TRPX_UMMA_TIMING")) {          // development aid: per-layer device time (synchronises)
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        cyl_umma_conv_kernel<L><<<grid, UM_THREADS, L::SMEM, st>>>(a);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        fprintf(stderr, "umma layer CIN=%d COUT=%d N=%d items=%d chunks/item=%d: %.1f us (%.0f cycles per chunk per CTA at 1.75 GHz)\n", L::CIN,
                L::COUT, a.N, a.nitems, L::KCH, ms * 1e3,
                ms * 1e-3 * 1.75e9 / ((double)((a.nitems + grid - 1) / grid) * L::KCH));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return TRPX_OK;
    }
    cyl_umma_conv_kernel<L><<<grid, UM_THREADS, L::SMEM, st>>>(a);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

// tcgen05 TF32 peak probe: every SM issues M128 x N256 x K8 MMAs (the math-bound shape) back to back from shared memory
__global__ void __launch_bounds__(128) umma_tf32_probe_kernel(int iters) {
    extern __shared__ __align__(128) unsigned char um_smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* z = reinterpret_cast<float*>(um_smem);
    for (int i = tid; i < (64 * 1024 + 32 * 1024) / 4; i += 128) z[i] = 0.f;
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) mbar_init(&bar, 1);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        constexpr uint32_t idesc = umma_idesc_tf32(128, 256);
        const uint32_t a0 = smem_u32(um_smem), b0 = a0 + 64 * 1024;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 16; j++)
                umma_tf32(tmem + (uint32_t)((j & 1) * 256), umma_desc(a0 + j * 4096, 2048), umma_desc(b0 + (j & 3) * 8192, 4096), idesc, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" int trpx_probe_umma_tf32(int device, float* tflops_out) {
    TRPX_REQUIRE(tflops_out != nullptr, "null output");
    int n = 0;
    TRPX_CUDA(cudaGetDeviceCount(&n));
    TRPX_REQUIRE(device >= 0 && device < n, "device %d out of range", device);
    DeviceGuard guard(device);
    int sms = 0;
    TRPX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int smem = 96 * 1024 + 128;
    TRPX_CUDA(cudaFuncSetAttribute(umma_tf32_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    TRPX_CUDA(cudaEventCreate(&e0));
    TRPX_CUDA(cudaEventCreate(&e1));
    const int iters = 2000;
    umma_tf32_probe_kernel<<<sms, 128, smem>>>(20);
    float best = 0.f;
    for (int rep = 0; rep < 3; rep++) {
        TRPX_CUDA(cudaEventRecord(e0));
        umma_tf32_probe_kernel<<<sms, 128, smem>>>(iters);
        TRPX_CUDA(cudaEventRecord(e1));
        TRPX_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        TRPX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 128 * 256 * 8 * 16.0 * (double)iters * (double)sms;
        best = std::max(best, (float)(flops / (ms * 1e-3) / 1e12));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    TRPX_CUDA(cudaGetLastError());
    *tflops_out = best;
    return TRPX_OK;
}

namespace trpx {

size_t umma_wimg_floats(int layer) {
    switch (layer) {
        case 0: return (size_t)UL0::KCH * UL0::W_BYTES / 4;
        case 1: return (size_t)UL1::KCH * UL1::W_BYTES / 4;
        case 2: return (size_t)UL2::KCH * UL2::W_BYTES / 4;
        default: return (size_t)UL3::KCH * UL3::W_BYTES / 4;
    }
}

int umma_build_wimg(int layer, const float* w, float* img, cudaStream_t st) {
    const int cin[4] = {128, 64, 32, 16}, cout[4] = {64, 32, 16, 16}, real[4] = {64, 32, 16, 3};
    TRPX_REQUIRE(layer >= 0 && layer < UMMA_LAYERS, "bad layer");
    umma_wimg_kernel<<<64, 256, 0, st>>>(w, img, cin[layer], cout[layer], real[layer]);
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

size_t umma_enc_wimg_floats(int layer) {
    switch (layer) {
        case 0: return (size_t)UE1::KCH * UE1::W_BYTES / 4;
        case 1: return (size_t)UE2::KCH * UE2::W_BYTES / 4;
        case 2: return 2 * (size_t)UE3::KCH * UE3::W_BYTES / 4;       // two groups of 32 output channels
        case 3: return 4 * (size_t)UE4::KCH * UE4::W_BYTES / 4;       // four groups of 32 output channels
        default: return (size_t)UE5::KCH * UE5::W_BYTES / 4;
    }
}

int umma_enc_build_wimg(int layer, const float* w, float* img, cudaStream_t st) {
    const int cin[4] = {4, 16, 32, 64}, cout[4] = {16, 32, 64, 128};
    TRPX_REQUIRE(layer >= 0 && layer < UMMA_ENC_LAYERS, "bad layer");
    if (layer == 4) {
        umma_wimg_kernel<<<64, 256, 0, st>>>(w, img, 128, 16, 4);
    } else if (layer == 3) {
        const size_t part = (size_t)UE4::KCH * UE4::W_BYTES / 4;
        for (int g = 0; g < 4; g++) umma_wimg_s2d_kernel<<<128, 256, 0, st>>>(w, img + g * part, 64, 32, 32 * g, 128);
    } else if (layer == 2) {
        const size_t part = (size_t)UE3::KCH * UE3::W_BYTES / 4;
        for (int g = 0; g < 2; g++) umma_wimg_s2d_kernel<<<128, 256, 0, st>>>(w, img + g * part, 32, 32, 32 * g, 64);
    } else {
        umma_wimg_s2d_kernel<<<64, 256, 0, st>>>(w, img, cin[layer], cout[layer], 0, cout[layer]);
    }
    TRPX_CUDA(cudaGetLastError());
    return TRPX_OK;
}

size_t umma_enc_ws_floats(int N) {
    // Q1 [4 quads][33][65], Q2 [16][17][33], Q3 [32][9][17], Q4 [64][5][9], a4 [32][4][8] (float4 each), z [128]
    return (size_t)N * (4 * (4 * 33 * 65 + 16 * 17 * 33 + 32 * 9 * 17 + 64 * 5 * 9 + 32 * 4 * 8) + 128);
}

int umma_encoder(const float* x, const float* visc, const float* mu, const float* sd, const float* const* wimg,
                 const float* const* bias, float* ws, float* z_out, int N, int sm_count, cudaStream_t st) {
    float* q1 = ws;
    float* q2 = q1 + (size_t)N * 4 * 4 * 33 * 65;
    float* q3 = q2 + (size_t)N * 4 * 16 * 17 * 33;
    float* q4 = q3 + (size_t)N * 4 * 32 * 9 * 17;
    float* a4 = q4 + (size_t)N * 4 * 64 * 5 * 9;
    const long long tot = (long long)N * 4 * 33 * 65;
    umma_enc_prep_kernel<<<(int)std::min<long long>((tot + 255) / 256, 148 * 16), 256, 0, st>>>(x, visc, mu, sd, reinterpret_cast<float4*>(q1), N);
    TRPX_CUDA(cudaGetLastError());
    UmmaArgs a{};
    a.N = N;
    int rc;
    a.in = q1; a.wimg = wimg[0]; a.bias = bias[0]; a.out = q2; a.co_base = 0; a.co_total = 16;
    if ((rc = launch_umma<UE1>(a, sm_count, st))) return rc;
    a.in = q2; a.wimg = wimg[1]; a.bias = bias[1]; a.out = q3; a.co_total = 32;
    if ((rc = launch_umma<UE2>(a, sm_count, st))) return rc;
    const size_t part3 = (size_t)UE3::KCH * UE3::W_BYTES / 4, part4 = (size_t)UE4::KCH * UE4::W_BYTES / 4;
    for (int g = 0; g < 2; g++) {
        a.in = q3; a.wimg = wimg[2] + g * part3; a.bias = bias[2] + g * 32; a.out = q4; a.co_base = g * 32; a.co_total = 64;
        if ((rc = launch_umma<UE3>(a, sm_count, st))) return rc;
    }
    for (int g = 0; g < 4; g++) {
        a.in = q4; a.wimg = wimg[3] + g * part4; a.bias = bias[3] + g * 32; a.out = a4; a.co_base = g * 32; a.co_total = 128;
        if ((rc = launch_umma<UE4>(a, sm_count, st))) return rc;
    }
    a.co_base = 0; a.co_total = 0;
    a.in = a4; a.wimg = wimg[4]; a.bias = bias[4]; a.out = z_out;
    return launch_umma<UE5>(a, sm_count, st);
}

int umma_conv_layer(int layer, const float* in_cl4, const float* wimg, const float* bias, float* out, int N,
                    const float* mu, const float* sd, const unsigned char* mask, int sm_count, cudaStream_t st) {
    UmmaArgs a{};
    a.in = in_cl4; a.wimg = wimg; a.bias = bias; a.out = out; a.mu = mu; a.sd = sd; a.mask = mask; a.N = N;
    static const int couts[4] = {64, 32, 16, 16};
    a.co_base = 0; a.co_total = couts[layer & 3];
    switch (layer) {
        case 0: return launch_umma<UL0>(a, sm_count, st);
        case 1: return launch_umma<UL1>(a, sm_count, st);
        case 2: return launch_umma<UL2>(a, sm_count, st);
        case 3: return launch_umma<UL3>(a, sm_count, st);
        default: return set_error(TRPX_E_INVALID, "bad layer %d", layer);
    }
}

}  // namespace trpx
