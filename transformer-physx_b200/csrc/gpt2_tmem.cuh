// Rollout with the KV window ON CHIP (tensor memory) for E = 32 / 4 heads / <= 4 layers: interface between the
// dispatcher in gpt2.cu and the kernel in gpt2_tmem.cu.
#pragma once
#include "common.cuh"
#include "gpt2_layout.cuh"

namespace trpx {

struct TmemRolloutParams {
    const float* blob_sw;     // XOR-swizzled weight image (swizzle_weights_kernel)
    const float* pe;
    const float* x0;
    long long x0_stride;
    float* out;
    long long out_stride;
    float* presents;          // optional [n_layer][2][B][H][Tk][hd]
    int B, S;
    int use_cache;            // 0: the reference's default generate() mode (no window, position 0)
    float eps;
    Gpt2Layout lay;
};

// trajectories one CTA keeps resident (8 for n_ctx = 32, 4 for n_ctx = 64); 0 if the shape is not supported
int tmem_rollout_group(const Gpt2Layout& L);
size_t tmem_rollout_smem(const Gpt2Layout& L);
// launches rollout_e32t_kernel<n_ctx>; returns a TRPX_* code
int launch_rollout_tmem(const TmemRolloutParams& p, int grid, cudaStream_t st);

}  // namespace trpx
