// Handle of the decoder stack, shared by gpt2.cu (rollout / forward) and gpt2_train.cu (training step).
#pragma once
#include "common.cuh"
#include "gpt2_layout.cuh"

struct trpx_gpt2 {
    uint32_t magic = 0x67707432u;
    trpx_gpt2_config cfg{};
    trpx::Gpt2Layout lay{};
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    float* blob = nullptr;
    float* blob_sw = nullptr;
    float* img_cl = nullptr;       // per-(image, rank) weight slices for the cluster kernel
    float* pe = nullptr;
    bool loaded = false;
    float* kv = nullptr;          // rollout workspace (KV windows of the groups in flight)
    size_t kv_bytes = 0;
    int tune_variant = 0, tune_warps = 0, tune_g = 0, tune_max_ctas = 0, tune_prefetch = 1;
    int last_variant = 0, last_grid = 0, last_block = 0, last_smem = 0, last_group = 0;
    float* train_ws = nullptr;    // training step: saved activations + per-sequence gradient partials
    size_t train_ws_bytes = 0;
};


namespace trpx {
inline bool valid(trpx_gpt2_t h) { return h != nullptr && h->magic == 0x67707432u; }
}  // namespace trpx
