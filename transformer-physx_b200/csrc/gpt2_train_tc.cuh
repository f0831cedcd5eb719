// Tensor-core training step for the E = 32 / 4-head decoder stacks (gpt2_train_tc.cu), called from trpx_gpt2_train_fwd_bwd.
#pragma once
#include "common.cuh"
#include "gpt2_layout.cuh"

namespace trpx {
bool train_tc_supported(const Gpt2Layout& L, int T);
size_t train_tc_acts_per_seq(const Gpt2Layout& L, int T);
// per-sequence partial gradients into gpart [B][L.total] and loss partials into loss_part [B]; train_finalize_kernel sums them
int launch_train_tc(const Gpt2Layout& L, const float* blob, const float* pe, const float* x, const float* labels, float* hidden,
                    float* acts, float* gpart, float* loss_part, int B, int T, float eps, cudaStream_t st);
}  // namespace trpx
