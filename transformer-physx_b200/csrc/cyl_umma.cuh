// tcgen05 (UMMA) implicit-GEMM layers of the Cylinder decoder (csrc/cyl_umma.cu), used by csrc/cyl.cu.
// Reference: trphysx/embedding/embedding_cylinder.py:70-88 (recoveryNet), :156-170 (recover).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace trpx {

// Decoder layers on the UMMA path.  Activations between them are "channels-last-4": [n][C/4][H][W][4] floats.
//   0: bilinear x2 (8x16 -> 16x32)  + conv 128 -> 64 + ReLU
//   1: bilinear x2 (16x32 -> 32x64) + conv  64 -> 32 + ReLU
//   2: bilinear x2 (32x64 -> 64x128)+ conv  32 -> 16 + ReLU
//   3: conv 16 -> 3 (64x128), un-normalise, cylinder mask; output planar [n][3][64][128]
constexpr int UMMA_LAYERS = 4;

// floats of the pre-arranged (TF32 hi / lo split, K-major core-matrix order) weight image of a layer
size_t umma_wimg_floats(int layer);
// w: [Cout][Cin][3][3] as in the state_dict; img: umma_wimg_floats(layer) floats
int umma_build_wimg(int layer, const float* w, float* img, cudaStream_t st);
// One layer over N frames.  mu / sd / mask are used by layer 3 only (mask may be null).
int umma_conv_layer(int layer, const float* in_cl4, const float* wimg, const float* bias, float* out, int N,
                    const float* mu, const float* sd, const unsigned char* mask, int sm_count, cudaStream_t st);

// ---- encoder (embedding_cylinder.py:42-60) on the same kernel -----------------------------------------------------
// The four stride-2 layers run in space-to-depth form (a stride-2 3x3 conv is a 2x2 conv over 2x2 pixel blocks with
// 4 x the channels), each layer's epilogue writing the next layer's block image directly; the fifth layer is a plain
// stride-1 conv.  Layers: 0: 4->16, 1: 16->32, 2: 32->64, 3: 64->128, 4: 128->4 (4x8).
constexpr int UMMA_ENC_LAYERS = 5;
size_t umma_enc_wimg_floats(int layer);
int umma_enc_build_wimg(int layer, const float* w, float* img, cudaStream_t st);
size_t umma_enc_ws_floats(int N);
// x [N][3][64][128], visc [N] -> z_out [N][128] (the encoder output before observableNetFC's LayerNorm)
int umma_encoder(const float* x, const float* visc, const float* mu, const float* sd, const float* const* wimg,
                 const float* const* bias, float* ws, float* z_out, int N, int sm_count, cudaStream_t st);

}  // namespace trpx
