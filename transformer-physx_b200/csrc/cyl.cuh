// Internal definitions shared by the Cylinder embedding sources (csrc/cyl.cu, csrc/cyl_train.cu).
#pragma once
#include "common.cuh"
#include "cyl_umma.cuh"

namespace trpx {
constexpr int CYL_H = 64, CYL_W = 128;
constexpr int NBAND = 4;
// tensor order of trpx_cyl_load_weights (include/trpx.h)
enum { T_MU = 0, T_STD = 1, T_ENC = 2, T_LNW = 12, T_LNB = 13, T_DEC = 14, T_KD = 24, T_KU = 28, T_KL = 32 };

// bilinear x2 (align_corners=True) source index / weight, as aten's area_pixel_compute_source_index
struct Lerp {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Lerp lerp_idx(int dst, int in_size, float scale) {
    // real = scale * dst ; i0 = min(int(real), in-1) ; l1 = clamp(real - i0, 0, 1) ; i1 = i0 + (i0 < in-1)
    const float real = scale * (float)dst;
    Lerp r;
    r.i0 = min((int)real, in_size - 1);
    r.l1 = fminf(fmaxf(real - (float)r.i0, 0.f), 1.f);
    r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
    r.l0 = 1.f - r.l1;
    return r;
}
}  // namespace trpx

struct trpx_cyl {
    uint32_t magic = 0x63796c31u;
    int E = 128;
    float eps = 1e-5f;
    int device = 0;
    int sm_count = 0;
    bool loaded = false;
    float* blob = nullptr;               // all 36 tensors back to back
    size_t off[36] = {0};
    size_t total = 0;
    unsigned char* mask = nullptr;       // [64][128] cylinder mask
    float* ws = nullptr;                 // activation workspace for one chunk of frames
    size_t ws_bytes = 0;
    int chunk = 64;                      // frames per decoder/encoder pass (intermediates stay in L2)
    int conv_mode = 3;                   // 3 = tcgen05 (UMMA) implicit GEMM, patch built in-kernel (default, csrc/cyl_umma.cu);
                                         // 2 = pipelined 3xTF32 mma.sync (upsample/split pass + cp.async double-buffered MMA
                                         // kernel), 1 = FP32 SIMT direct conv, 0 = 3xTF32 mma.sync with in-kernel patch build
    int final_mma = 0;                   // route the last 16->3 conv through the mode-2 pipeline as well (slower, kept for A/B)
    bool chunk_user = false;
    float* wimg = nullptr;               // pre-split weight images of the three heavy decoder layers (mode 2)
    size_t wimg_off[4] = {0, 0, 0, 0};
    float* U = nullptr;                  // upsampled / padded / split layer input for one chunk (mode 2)
    size_t U_bytes = 0;
    float* uimg = nullptr;               // weight images of the four UMMA layers (mode 3)
    size_t uimg_off[trpx::UMMA_LAYERS + 1] = {0};
    float* eimg = nullptr;               // weight images of the five UMMA encoder layers (mode 3)
    size_t eimg_off[trpx::UMMA_ENC_LAYERS + 1] = {0};
    float* tws = nullptr;                // training-step workspace (csrc/cyl_train.cu)
    size_t tws_bytes = 0;
};

namespace trpx {
inline bool cyl_valid(trpx_cyl_t h) { return h != nullptr && h->magic == 0x63796c31u; }
// element counts of the 36 tensors, in the order documented in include/trpx.h
void cyl_tensor_sizes(int E, size_t (&n)[36]);
}  // namespace trpx
