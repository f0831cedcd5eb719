// tcgen05 (UMMA) recover kernel for the Lorenz / Rossler embedding (mlpemb_umma.cu), used by trpx_mlpemb_recover for large N.
#pragma once
#include "mlpemb.cuh"

namespace trpx {
bool mlp_recover_umma_ok(trpx_mlpemb_t h, long long N);
int mlp_recover_umma_grid(trpx_mlpemb_t h, long long N);
int mlp_recover_umma(trpx_mlpemb_t h, const float* g, float* x, long long N, const float* target, float* partial,
                     cudaStream_t st);
}  // namespace trpx
