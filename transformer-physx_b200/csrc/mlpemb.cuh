// Lorenz / Rossler Koopman embedding: handle + packed weight layout shared by mlpemb.cu and mlpemb_train.cu.
// Reference: trphysx/embedding/embedding_lorenz.py:33-71 (module shapes), :94-142 (embed/recover/Koopman).
#pragma once
#include "common.cuh"

struct trpx_mlpemb {
    uint32_t magic = 0x6d6c7065u;
    trpx_mlpemb_config cfg{};
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    bool loaded = false;
    float* blob = nullptr;        // packed weights, see MlpLayout
    // training workspace (mlpemb_train.cu)
    float* ws = nullptr;
    size_t ws_bytes = 0;
    float* mse_partial = nullptr; // per-CTA squared-error sums of the fused recover + MSE
    int mse_partial_n = 0;
    int recover_mode = 0;         // 0 automatic, 1 FP32 SIMT kernel only (A/B)
};

namespace trpx {

constexpr int ME = 32;   // n_embd
constexpr int MS = 3;    // state dim

// Packed layout (floats).  Hd = hidden width (500).
//   encA  [Hd][4]   {W1[u][0], W1[u][1], W1[u][2], b1[u]}            observableNet.0
//   encB  [Hd][32]  W2[e][u] transposed to [u][e]                      observableNet.2.weight^T
//   b2    [32], lnw [32], lnb [32]                                     observableNet.2.bias, .3.{weight,bias}
//   decA  [Hd][32]  V1[u][k]                                           recoveryNet.0.weight
//   decB  [Hd][4]   {V2[0][u], V2[1][u], V2[2][u], c1[u]}            recoveryNet.2.weight^T | recoveryNet.0.bias
//   c2    [4]       recoveryNet.2.bias (padded)
//   mu [4], sd [4]  buffers (padded)
//   kd [32], ku [64] kMatrixDiag, kMatrixUT (61, padded)
// Original (unpacked) copies needed by the backward pass follow: W2 [32][Hd], V2 [3][Hd].
struct MlpLayout {
    int Hd;
    int encA, encB, b2, lnw, lnb, decA, decB, c2, mu, sd, kd, ku, W2, V2, total;
    __host__ __device__ static MlpLayout make(int Hd) {
        MlpLayout L;
        L.Hd = Hd;
        int o = 0;
        L.encA = o; o += Hd * 4;
        L.encB = o; o += Hd * ME;
        L.b2 = o; o += ME;
        L.lnw = o; o += ME;
        L.lnb = o; o += ME;
        L.decA = o; o += Hd * ME;
        L.decB = o; o += Hd * 4;
        L.c2 = o; o += 4;
        L.mu = o; o += 4;
        L.sd = o; o += 4;
        L.kd = o; o += ME;
        L.ku = o; o += 64;
        L.W2 = o; o += ME * Hd;
        L.V2 = o; o += MS * Hd;
        L.total = o;
        return L;
    }
};

// embed that can also save the pre-affine LayerNorm activation and 1/std (used by the training step)
int mlpemb_embed_ex(trpx_mlpemb_t h, const float* x, float* g, long long N, float* zhat, float* rstd, void* stream);

inline bool valid(trpx_mlpemb_t h) { return h != nullptr && h->magic == 0x6d6c7065u; }

}  // namespace trpx
