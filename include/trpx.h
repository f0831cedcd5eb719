/*
 * trpx.h — C ABI of the B200-native rollout path of transformer-physx (libtrpx.so).
 *
 * The reference (zabaras/transformer-physx, `trphysx` 0.0.8) is pure Python/PyTorch and has no FFI
 * layer of its own; the boundary below is what a binding for its rollout hot path would call.  Each
 * entry point names the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer named *_dev / documented "device" is a CUDA device pointer on the handle's device;
 *     all tensors are float32, row-major contiguous unless a stride argument says otherwise;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls only ENQUEUE work;
 *   - return value 0 = OK, non-zero = error (TRPX_E_*); trpx_last_error() returns the message of the
 *     last failing call on the calling thread.  No C++ exception crosses this boundary;
 *   - a handle is bound to one device, is not re-entrant, and owns only packed weights + workspace
 *     (KV windows, scratch).  Inputs/outputs are always caller-owned.
 *   - there is NO CPU implementation behind any of these calls.
 */
#ifndef TRPX_H
#define TRPX_H

#ifdef __cplusplus
extern "C" {
#endif

#define TRPX_OK              0
#define TRPX_E_INVALID       1   /* bad argument / unsupported configuration */
#define TRPX_E_CUDA          2   /* a CUDA runtime call failed */
#define TRPX_E_STATE         3   /* weights not loaded, handle destroyed, ... */

const char* trpx_last_error(void);
int         trpx_version(void);                 /* 100*major + minor */
/* SM count / max opt-in shared memory of `device` (for tests and the bench) */
int         trpx_device_info(int device, int* sm_count, int* smem_optin_bytes);
/* Measured FP32 FMA throughput of `device` in TFLOP/s (register-only FMA chains; the roofline denominator of
 * the FP32-bound kernels, which MEASURED_PEAKS.json does not carry).  Synchronous. */
int         trpx_probe_fp32_fma(int device, float* tflops_out);
/* Measured dense TF32 throughput of the warp-level mma.sync.m16n8k8 path (SASS HMMA.1688.F32.TF32) in TFLOP/s:
 * the denominator (divided by 3 for the error-compensated 3xTF32 product) of the tensor-core rollout kernel. */
int         trpx_probe_tf32_mma(int device, float* tflops_out);
/* Dense TF32 rate of the 5th-generation tensor cores (tcgen05.mma kind::tf32, M128 x N256 x K8 from shared memory,
 * accumulators in TMEM) in TFLOP/s: the denominator of the Cylinder decoder's roofline (its error-compensated 3xTF32
 * product runs at a third of it). */
int         trpx_probe_umma_tf32(int device, float* tflops_out);
/* Measured tcgen05.ld rate (bytes per clock per SM, 8 warps x 32x32b.x32 loads): what bounds the window read of the
 * rollout kernel that keeps the KV window in tensor memory (rollout variant 10, csrc/gpt2_tmem.cu). */
int         trpx_probe_tmem_ld(int device, float* bytes_per_clk_out);
/* Measured rate (GB/s) of the KV-window access pattern of the cache-mode rollout with the arithmetic removed:
 * 4 trajectories per warp, per (step, layer) two windows of n_ctx rows of 128 B read with 256-bit loads
 * (`rows_in_flight` per lane: 4, 8 or 16) plus two rows appended; `prefetch` adds the bulk L2 prefetch the rollout
 * issues.  It is the practical ceiling of the HBM-bound rollout kernel for this access pattern.  Synchronous. */
int         trpx_probe_window_stream(int device, int n_traj, int n_ctx, int n_layer, int steps, int warps_per_cta,
                                     int rows_in_flight, int prefetch, float* gbps_out, float* ms_out);

/* ------------------------------------------------------------------------------------------------
 * Transformer decoder stack: PhysformerGPT2 (trphysx/transformer/phys_transformer_gpt2.py:120-269)
 * ---------------------------------------------------------------------------------------------- */
typedef struct trpx_gpt2* trpx_gpt2_t;

typedef struct {
    int   n_ctx;     /* PhysConfig.n_ctx   (configuration_phys.py:54) */
    int   n_embd;    /* PhysConfig.n_embd  (:55)  multiple of 32, <= 512 (Gray-Scott: 512) */
    int   n_layer;   /* PhysConfig.n_layer (:56) */
    int   n_head;    /* PhysConfig.n_head  (:57)  n_embd % n_head == 0 */
    float ln_eps;    /* PhysConfig.layer_norm_epsilon (:68) */
} trpx_gpt2_config;

/* PhysformerGPT2.__init__ (phys_transformer_gpt2.py:126-146): allocates the packed-weight blob. */
int trpx_gpt2_create(const trpx_gpt2_config* cfg, int device, trpx_gpt2_t* out);
int trpx_gpt2_destroy(trpx_gpt2_t h);

/* Number of weight tensors trpx_gpt2_load_weights expects: 12*n_layer + 4. */
int trpx_gpt2_num_weight_tensors(trpx_gpt2_t h);

/* load_state_dict (phys_transformer_base.py:120-137).  `w_dev[i]` are device pointers, in this order:
 *   for l in 0..n_layer-1:  h.l.ln_1.weight[E] ln_1.bias[E] attn.c_attn.weight[E,3E] attn.c_attn.bias[3E]
 *                           attn.c_proj.weight[E,E] attn.c_proj.bias[E] ln_2.weight[E] ln_2.bias[E]
 *                           mlp.c_fc.weight[E,4E] mlp.c_fc.bias[4E] mlp.c_proj.weight[4E,E] mlp.c_proj.bias[E]
 *   then ln_f.weight[E] ln_f.bias[E] mlp_f.weight[E_out,E_in] mlp_f.bias[E]
 * Conv1D weights are [in,out] (utils.py:37-39); mlp_f is nn.Linear [out,in] and is transposed here.
 * `pe_table_dev` = position embedding [n_ctx, E] (phys_transformer_gpt2.py:215-219) or NULL to have the
 * library compute it (even channels sin(p/10000^(2j/E)), odd channels cos(p)). */
int trpx_gpt2_load_weights(trpx_gpt2_t h, const float* const* w_dev, int n_tensors,
                           const float* pe_table_dev, void* stream);

/* GenerationMixin.generate / _generate_time_series (generate_utils.py:66-169).
 *   x0_dev            last context token of each trajectory, element (b, e) at x0_dev[b*x0_stride + e]
 *   out_dev           generated tokens, element (b, s, e) at out_dev[b*out_stride + s*E + e], s in [0,S)
 *   use_cache != 0    sliding KV window of n_ctx keys, position id min(s, n_ctx-1)  (generate_utils.py:156-157)
 *   use_cache == 0    the reference default: every step is a 1-token forward at position 0
 *   presents_dev      NULL, or (use_cache only) [n_layer][2][B][n_head][Tk][hd] with Tk = min(S, n_ctx):
 *                     the `presents` of the LAST step in chronological order (attention.py:188)
 * One launch covers the whole horizon; trajectories are independent. */
int trpx_gpt2_rollout(trpx_gpt2_t h, const float* x0_dev, long long x0_stride,
                      float* out_dev, long long out_stride, int B, int S, int use_cache,
                      float* presents_dev, void* stream);

/* Teacher-forced training step (replaces PhysformerTrain.forward + loss.backward(),
 * trphysx/transformer/phys_transformer_helpers.py:36-69 and utils/trainer.py:244-277):
 *   hidden = forward(x)[0];  loss = mean((hidden - labels)^2) over B*T*n_embd
 * x, labels: [B][T][n_embd], T <= n_ctx, positions 0..T-1, no past.  Writes the scalar loss, optionally the hidden
 * states ([B][T][n_embd], may be NULL) and d loss / d weights as ONE blob of trpx_gpt2_blob_size() floats laid out
 * like the packed weights: per layer ln_1.w, ln_1.b, c_attn.w [E][3E], c_attn.b, c_proj.w [E][E], c_proj.b, ln_2.w,
 * ln_2.b, c_fc.w [E][4E], c_fc.b, mlp.c_proj.w [4E][E], mlp.c_proj.b; then ln_f.w, ln_f.b, mlp_f.weight TRANSPOSED
 * ([in][out]), mlp_f.b.  Deterministic (per-sequence partials summed in a fixed order). */
int trpx_gpt2_train_fwd_bwd(trpx_gpt2_t h, const float* x_dev, const float* labels_dev, int B, int T, float* loss_dev,
                            float* hidden_dev, float* grad_blob_dev, void* stream);
int trpx_gpt2_blob_size(trpx_gpt2_t h);
/* Replace the weights by a packed blob in the same layout (device pointer, trpx_gpt2_blob_size() floats): used by the
 * fused optimizer, whose master parameters live in this layout, after every clip + Adam update. */
int trpx_gpt2_load_blob(trpx_gpt2_t h, const float* blob_dev, void* stream);

/* Tuning knobs of the rollout kernel (0 = automatic).  `variant`: 0 auto, 1 generic kernel only, 2 E=32 warp
 * kernel, 3 = 2 without TMA weight staging, 4 = 2 without the L2 window prefetch (for A/B measurements),
 * 5 = E=32 kernel with 2-D lane tiling (16 trajectories per warp), 6 = E=32 tensor-core kernel (3xTF32
 * mma.sync; chosen automatically for large batches without the KV window), 7 = thread-block-cluster kernel
 * (8 CTAs per trajectory group, column-split GEMVs, DSMEM exchange, TMA weight ring; chosen automatically for
 * shapes other than E=32 when one trajectory per cluster fits in a single wave; groups of 1-2 trajectories exchange
 * with st.async + mbarrier complete_tx instead of cluster barriers), 9 = 7 with the cluster-barrier exchange (A/B),
 * 8 = E=32 window kernel with TMA-staged attention and a per-layer weight ring, 10 = E=32 kernel with the KV window on
 * chip in tensor memory (chosen automatically for small batches), 11 = warp-specialised E=32 window kernel (dense
 * tensor-core warps + attention warps; KV-window mode only, A/B).  With warps_per_cta = 0 the window kernel takes the
 * fewest warps per CTA that need no more rounds over the batch than the maximum (8) would. */
int trpx_gpt2_set_rollout_tuning(trpx_gpt2_t h, int variant, int warps_per_cta, int traj_per_warp,
                                 int max_ctas);
/* What the last trpx_gpt2_rollout launched: variant (1 generic, 2 E=32 warp kernel, 5 E=32 2-D tiling, 6 E=32 tensor-core,
 * 7 cluster kernel), grid, block,
 * dynamic smem bytes, trajectories per warp/CTA group. */
int trpx_gpt2_last_launch(trpx_gpt2_t h, int* variant, int* grid, int* block, int* smem, int* group);

/* PhysformerGPT2.forward (phys_transformer_gpt2.py:147-269) with position_ids/prop_embeds/
 * attention_mask/head_mask = None.
 *   x_dev         [B,T,E]
 *   past_dev      NULL or n_layer device pointers to contiguous [2,B,H,Tp,hd]     (attention.py:182-185)
 *   hidden_dev    [B,T,E]    mlp_f(ln_f(h))
 *   presents_dev  NULL or n_layer device pointers to [2,B,H,Tp+T,hd]              (attention.py:188)
 *   attn_dev      NULL or n_layer device pointers to [B,H,T,Tp+T] softmax weights (attention.py:109)
 * Requires Tp + T <= n_ctx (the causal-mask buffer is n_ctx x n_ctx, attention.py:101-102). */
int trpx_gpt2_forward(trpx_gpt2_t h, const float* x_dev, const float* const* past_dev, int Tp,
                      float* hidden_dev, float* const* presents_dev, float* const* attn_dev,
                      int B, int T, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Lorenz / Rossler Koopman embedding (trphysx/embedding/embedding_lorenz.py:25-167,
 * examples/rossler/rossler_module/embedding_rossler.py:23-161)
 * ---------------------------------------------------------------------------------------------- */
typedef struct trpx_mlpemb* trpx_mlpemb_t;

typedef struct {
    int   state_dim;   /* config.state_dims[0] = 3 */
    int   hidden;      /* hard-coded 500 (embedding_lorenz.py:39) */
    int   n_embd;      /* 32 */
    float ln_eps;
} trpx_mlpemb_config;

int trpx_mlpemb_create(const trpx_mlpemb_config* cfg, int device, trpx_mlpemb_t* out);
int trpx_mlpemb_destroy(trpx_mlpemb_t h);
/* 14 device pointers: mu[3] std[3] observableNet.0.weight[500,3] .0.bias[500] observableNet.2.weight[32,500]
 * .2.bias[32] observableNet.3.weight[32] .3.bias[32] recoveryNet.0.weight[500,32] .0.bias[500]
 * recoveryNet.2.weight[3,500] .2.bias[3] kMatrixDiag[32] kMatrixUT[61]   (all nn.Linear [out,in]) */
int trpx_mlpemb_load_weights(trpx_mlpemb_t h, const float* const* w_dev, int n_tensors, void* stream);
/* LorenzEmbedding.embed (embedding_lorenz.py:94-105): x[N,3] -> g[N,32] */
int trpx_mlpemb_embed(trpx_mlpemb_t h, const float* x_dev, float* g_dev, long long N, void* stream);
/* LorenzEmbedding.recover (:107-118): g[N,32] -> x[N,3] */
int trpx_mlpemb_recover(trpx_mlpemb_t h, const float* g_dev, float* x_dev, long long N, void* stream);
/* recover kernel selection: 0 automatic (tcgen05 kind::tf32 kernel, csrc/mlpemb_umma.cu, for N >= 4 x 128 x SM count;
 * 3xTF32 mma.sync kernel for N >= 16384; FP32 SIMT kernel below), 1 FP32 SIMT kernel only (A/B measurements,
 * bit-stable results), 2 never the tcgen05 kernel (A/B), 3 same as 0. */
int trpx_mlpemb_set_recover_mode(trpx_mlpemb_t h, int mode);
/* Fused `recover` + MSE of Trainer.eval_states (trphysx/utils/trainer.py:380-390):
 * mse = mean((recover(g) - target)^2) over N*3 values, written to mse_dev (device scalar).  The recovered states
 * are written only if states_or_null is given ([N][3]).  Deterministic (per-CTA sums added in a fixed order). */
int trpx_mlpemb_recover_mse(trpx_mlpemb_t h, const float* g_dev, const float* target_dev, long long N,
                            float* states_or_null, float* mse_dev, void* stream);
/* LorenzEmbedding.koopmanOperation (:120-142) as a banded matvec: g[N,32] -> g'[N,32] */
int trpx_mlpemb_koopman(trpx_mlpemb_t h, const float* g_dev, float* gnext_dev, long long N, void* stream);
/* LorenzEmbedding.koopmanOperator (:144-157): dense K[32,32] */
int trpx_mlpemb_koopman_matrix(trpx_mlpemb_t h, float* K_dev, void* stream);

/* LorenzEmbeddingTrainer.forward + backward (embedding_lorenz.py:181-221; enn_trainer.py:93-96).
 *   states_dev [B,T,3];  w_recon 1e4 (Lorenz) / 1e3 (Rossler); w_dyn 1; w_reg 1e-1
 *   loss2_dev[2] = {loss, loss_reconstruct};  grad_dev = flat gradient [36192] in nn.Module.parameters()
 *   order: kMatrixDiag, kMatrixUT, observableNet.{0.w,0.b,2.w,2.b,3.w,3.b}, recoveryNet.{0.w,0.b,2.w,2.b}.
 *   `grad_scale` multiplies the gradient (1/world_size before a SUM all-reduce). */
int trpx_mlpemb_train_fwd_bwd(trpx_mlpemb_t h, const float* states_dev, int B, int T,
                              float w_recon, float w_dyn, float w_reg, float grad_scale,
                              float* loss2_dev, float* grad_dev, void* stream);
long long trpx_mlpemb_num_params(trpx_mlpemb_t h);

/* clip_grad_norm_(max_norm) + torch.optim.Adam step on flat buffers (enn_trainer.py:97-99):
 * param/grad/m/v [n] device; step >= 1; norm_out_dev[1] receives the pre-clip global L2 norm. */
int trpx_clip_adam(float* param_dev, const float* grad_dev, float* m_dev, float* v_dev, long long n,
                   int step, float lr, float beta1, float beta2, float eps, float weight_decay,
                   float max_norm, float* norm_out_dev, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cylinder Koopman embedding (trphysx/embedding/embedding_cylinder.py:25-222)
 * ---------------------------------------------------------------------------------------------- */
typedef struct trpx_cyl* trpx_cyl_t;

int trpx_cyl_create(int n_embd /*128*/, float ln_eps, int device, trpx_cyl_t* out);
int trpx_cyl_destroy(trpx_cyl_t h);
/* 36 device pointers: mu[4] std[4]; observableNet.{0,2,4,6,8}.{weight,bias}; observableNetFC.0.{weight,bias};
 * recoveryNet.{1,4,7,10,12}.{weight,bias}; kMatrixDiagNet.{0,2}.{weight,bias}; kMatrixUT.{0,2}.{weight,bias};
 * kMatrixLT.{0,2}.{weight,bias}.  Conv weights are [Cout,Cin,3,3]. */
int trpx_cyl_load_weights(trpx_cyl_t h, const float* const* w_dev, int n_tensors, void* stream);
/* Frames per encoder/decoder pass (default 64): intermediates of one pass stay L2-resident. */
int trpx_cyl_set_chunk(trpx_cyl_t h, int frames);
/* Decoder conv path (embedding_cylinder.py:70-88): 3 = tcgen05 (UMMA) TF32 implicit GEMMs with TMA-staged layer
 * input and weights, the bilinear upsample + TF32 hi/lo split fused into the kernel's producer warps, accumulators in
 * TMEM (default, csrc/cyl_umma.cu); 2 = pipelined 3xTF32 mma.sync path (separate upsample/pad/split pass, cp.async
 * double-buffered MMA kernel); 1 = FP32 SIMT direct convolution; 0 = 3xTF32 mma.sync with in-kernel patch build.
 * 0-2 are kept for A/B measurements (profiles/README.md). */
int trpx_cyl_set_conv_mode(trpx_cyl_t h, int mode);
/* CylinderEmbedding.embed (:138-154): x[N,3,64,128], visc[N] -> g[N,128] */
int trpx_cyl_embed(trpx_cyl_t h, const float* x_dev, const float* visc_dev, float* g_dev, int N, void* stream);
/* CylinderEmbedding.recover (:156-170): g[N,128] -> x[N,3,64,128]; apply_mask=1 zeroes the cylinder
 * cells (recover) and 0 leaves them (the decode half of forward(), :128-134). */
int trpx_cyl_recover(trpx_cyl_t h, const float* g_dev, float* x_dev, int N, int apply_mask, void* stream);
/* CylinderEmbedding.koopmanOperation (:172-196): g[N,128], visc[N] -> g'[N,128];
 * kdiag_dev NULL or [N,128] (self.kMatrixDiag); kmat_dev NULL or [N,128,128] dense operator (self.kMatrix) */
int trpx_cyl_koopman(trpx_cyl_t h, const float* g_dev, const float* visc_dev, float* gnext_dev,
                     float* kdiag_dev, float* kmat_dev, int N, void* stream);

/* CylinderEmbeddingTrainer.forward + backward (embedding_cylinder.py:236-281; the step body of
 * embedding/training/enn_trainer.py:93-96) in one call:
 *   states_dev [B,T,3,64,128], visc_dev [B];  w_rec 1e1, w_dyn 1e1, w_reg 1e-2 in the reference
 *   loss2_dev[2] = {loss, loss_reconstruct};  grad_dev = flat gradient [trpx_cyl_num_params = 262,535] in
 *   nn.Module.parameters() order: observableNet.{0,2,4,6,8}.{weight,bias}, observableNetFC.0.{weight,bias},
 *   recoveryNet.{1,4,7,10,12}.{weight,bias}, kMatrixDiagNet / kMatrixUT / kMatrixLT .{0.weight,0.bias,2.weight,2.bias}.
 *   `grad_scale` multiplies the gradient (1/world_size before a SUM all-reduce).  Deterministic (no atomics). */
int trpx_cyl_train_fwd_bwd(trpx_cyl_t h, const float* states_dev, const float* visc_dev, int B, int T,
                           float w_rec, float w_dyn, float w_reg, float grad_scale,
                           float* loss2_dev, float* grad_dev, void* stream);
long long trpx_cyl_num_params(trpx_cyl_t h);

/* ------------------------------------------------------------------------------------------------
 * Gray-Scott Koopman embedding (trphysx/embedding/embedding_grayscott.py:25-218), eval mode: BatchNorm3d uses its
 * running statistics (folded into the convs at load time).  States are [N, 2, 32, 32, 32].
 * ---------------------------------------------------------------------------------------------- */
typedef struct trpx_gs* trpx_gs_t;

int trpx_gs_create(int n_embd /*512*/, float ln_eps, int device, trpx_gs_t* out);
int trpx_gs_destroy(trpx_gs_t h);
int trpx_gs_num_weight_tensors(trpx_gs_t h);   /* 64 */
/* 64 device pointers: mu[1] std[1]; for observableNet.{0,3,6,9} then recoveryNet.{0,4,8,12}: conv.{weight,bias} and the
 * BatchNorm3d that follows it {weight,bias,running_mean,running_var}; observableNetFC.{0,2}.{weight,bias},
 * observableNetFC.3.{weight,bias} (LayerNorm); recoveryNetFC.{0,2}.{weight,bias}; recoveryNet.15.{weight,bias};
 * kMatrixDiag[E]; kMatrixUT[sum_{d=1..9}(E-d)].  Conv weights are [Cout,Cin,k,k,k], Linear weights [out,in]. */
int trpx_gs_load_weights(trpx_gs_t h, const float* const* w_dev, int n_tensors, void* stream);
/* GrayScottEmbedding.embed (:127-139): x[N,2,32,32,32] -> g[N,E]; g0_dev NULL or [N,64,4,4,4] = observableNet(x),
 * the third element of forward()'s tuple (:112-125) */
int trpx_gs_embed(trpx_gs_t h, const float* x_dev, float* g_dev, float* g0_dev, int N, void* stream);
/* GrayScottEmbedding.recover (:141-153): g[N,E] -> x[N,2,32,32,32]; out0_dev NULL or [N,64,4,4,4] = recoveryNetFC(g) */
int trpx_gs_recover(trpx_gs_t h, const float* g_dev, float* x_dev, float* out0_dev, int N, void* stream);
/* unnormalize(recoveryNet(z)) for a latent volume z[N,64,4,4,4] (last element of forward()'s tuple) */
int trpx_gs_decode_latent(trpx_gs_t h, const float* z_dev, float* x_dev, int N, void* stream);
/* GrayScottEmbedding.koopmanOperation (:155-175): g[B,E] -> g'[B,E]; kmat_dev NULL or [B,E,E] (self.kMatrix) */
int trpx_gs_koopman(trpx_gs_t h, const float* g_dev, float* gnext_dev, float* kmat_dev, int B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRPX_H */
