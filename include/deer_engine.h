/*
 * deer_engine.h -- C ABI of the B200-native DeeR-VLA multi-exit inference engine.
 *
 * The reference (yueyang130/DeeR-VLA) is pure Python: it has NO plugin / FFI interface.
 * Its drop-in boundary for the per-timestep hot path is the Python call surface
 *     ModelWrapper.step            robot_flamingo/eval/eval_utils.py:279-490
 *     MPTFlamingo.forward          robot_flamingo/models/flamingo_mpt.py:308-461
 *     ExitController.forward       robot_flamingo/models/value_net.py:277-297
 * Each entry point below names the reference interface it replaces.  Plain pointers and
 * sizes only; no torch / C++ types; never throws.  All functions return 0 on success and
 * a negative DEER_E_* code on failure; deer_last_error() gives the message.
 *
 * Threading: one ctx per GPU; calls on one ctx must be serialised by the caller (the
 * reference module is not re-entrant either: flamingo_lm.py:37-44, action_head.py:474).
 */
#ifndef DEER_ENGINE_H
#define DEER_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEER_ABI_VERSION 2   /* 2: deer_config grew mlp_dim2 / mlp_layernorm / lstm_layernorm / action_ensemble */

enum {
  DEER_OK = 0,
  DEER_E_INVALID = -1,   /* bad argument (mirrors the reference's python asserts) */
  DEER_E_CUDA = -2,      /* a CUDA call failed */
  DEER_E_STATE = -3,     /* call order (weights not finalized, thresholds not set ...) */
  DEER_E_MISSING = -4    /* a required weight was never loaded */
};

enum { DEER_PREC_BF16 = 0,   /* product path: bf16 operands, tcgen05 tensor cores, fp32 accumulate */
       DEER_PREC_FP32 = 1 }; /* validation path: fp32 everywhere, CUDA-core GEMMs (slow) */

enum { DEER_GEMM_TCGEN05 = 0, DEER_GEMM_SIMT = 1,      /* SIMT = debug fallback selector for bf16 */
       DEER_GEMM_TC_1CTA = 2, DEER_GEMM_TC_SPLITK = 3, DEER_GEMM_TC_1CTA_M128 = 4 };   /* deer_op_gemm A/B selectors: no 2-CTA kernel /
       128-row tiles with split-K for few-row launches / 128-row tiles only (no 64 x 64 tiles, no split-K) */

enum { DEER_THR_L2 = 0, DEER_THR_MEAN = 1, DEER_THR_MAX = 2, DEER_THR_COSINE = 3 }; /* value_net.py:105-117 */

/* deer_step flags */
#define DEER_FLAG_FULL_DEPTH 1u  /* calibration: ignore thresholds, run to the last enabled exit,
                                    record delta at every exit (value_net.py:300-399 "generate" use) */

#define DEER_FLAG_GENERATE 2u    /* internal to deer_step_generate (implies FULL_DEPTH) */

typedef struct deer_ctx deer_ctx;

/* Model + exit configuration.  Mirrors what eval_calvin.py derives from its flags and the
 * checkpoint (eval_calvin.py:455-540) and passes to create_model_and_transforms
 * (robot_flamingo/models/factory.py:53-269). */
typedef struct deer_config {
  int32_t abi_version;       /* must be DEER_ABI_VERSION */
  /* vision tower: open_clip ViT (factory.py:109-114) */
  int32_t image_size, patch, vit_width, vit_layers, vit_heads, vit_mlp;
  /* perceiver resampler (open_flamingo/src/helpers.py:68-81) */
  int32_t perc_depth, perc_heads, perc_dim_head, num_latents, perc_ff_mult;
  /* language model (MPT-1B for 3B flamingo, MPT-7B for 9B flamingo; factory.py:13-26) */
  int32_t d_model, n_heads, mlp_ratio, vocab;
  int32_t qk_ln;             /* 1: MPT-1B GPTBlock (attn_qk_ln, ln_1/ln_2/mlp_up/mlp_down names);
                                0: MPT-7B MPTBlock (norm_1/norm_2/ffn.up_proj/ffn.down_proj) */
  int32_t alibi_bias_max;
  int32_t cross_attn_every;  /* flamingo_lm.py:174-176 */
  int32_t xattn_heads, xattn_dim_head, xattn_ff_mult;   /* helpers.py:236-258 */
  int32_t media_token_id;    /* flamingo_lm.py:211 */
  /* exits */
  int32_t n_layers;          /* layers built = min(ckpt early_exit_layer, max_layer)+1 (eval_calvin.py:529, flamingo_mpt.py:198) */
  int32_t exit_interval;     /* flamingo_mpt.py:248-250 */
  int32_t max_layer;         /* --max_layer as given on the CLI (value_net.py:173 derives min(max_layer-1, E[-1])) */
  int32_t is_mpt_9b;         /* value_net.py:235-236 (only used by the host-side threshold solver) */
  /* action head (action_head.py:408-491) */
  int32_t lstm_hidden, lstm_layers, mlp_dim0, mlp_dim1;
  int32_t threshold_type;    /* DEER_THR_* */
  int32_t steps_per_stage;   /* value_net.py:285-286 */
  /* engine sizing */
  int32_t max_batch;         /* max concurrent episodes per deer_step call (<= 1024) */
  int32_t max_slots;         /* number of per-episode LSTM state slots the engine owns */
  int32_t max_seq;           /* max text tokens S (reference: 32, data.py:912-918) */
  int32_t precision;         /* DEER_PREC_* */
  int32_t gemm_impl;         /* DEER_GEMM_* (bf16 only) */
  int32_t debug_taps;        /* 1: keep per-stage copies readable through deer_get_tap */
  int32_t host_control;      /* 1: read the active count back after every probe (the reference's
                                per-probe D2H sync, value_net.py:293) and stop launching when 0;
                                0: fully device-side control flow, no host sync inside a step */
  int32_t use_graph;         /* 1 (needs host_control=0): each (n, S, mode) step is captured ONCE into a CUDA graph whose
                                post-exit segments are nested conditional IF nodes; exit_gate_kernel sets the condition
                                (cudaGraphSetConditional) from the device-side active count, so the layers past the
                                last active episode's exit are never launched and no host round trip is needed */
  /* head variants of older DeeR checkpoints (eval_calvin.py:464-476; action_head.py:66-79, 82-161): the trained DeeR
   * configuration is mlp_dim2 = 0 (two hidden layers), mlp_layernorm = lstm_layernorm = 1 */
  int32_t mlp_dim2;          /* > 0: a third hidden layer of this width (mlp_num_hidden_layers=3: 1024, 512, 256) */
  int32_t mlp_layernorm;     /* LayerNorm after every hidden Linear of both MLP heads */
  int32_t lstm_layernorm;    /* 1: LayerNormLSTM (stored h pre-LN, next layer's input post-LN); 0: plain stacked nn.LSTM */
  int32_t action_ensemble;   /* 1: action_out = mean of the exit's and the previous exit's (pose, gripper) -- ModelWrapper's
                                use_action_ensemble (eval_utils.py:457-462, value_net.py:92-95); pose_out / gripper_out
                                stay the exit head's own outputs */
} deer_config;

typedef struct deer_stats {
  int64_t kernel_launches_total;      /* kernels of this library launched since create */
  int64_t kernel_launches_last_step;  /* ... by the most recent deer_step / deer_step_host */
  int64_t gemm_launches_last_step;
  double  gemm_flops_last_step;       /* 2*M*N*K summed over GEMM launches with host-known M upper bound */
  int64_t steps;
  int64_t weight_bytes;               /* device bytes held by weights */
  int64_t workspace_bytes;
} deer_stats;

/* library-level */
int         deer_abi_version(void);
const char* deer_build_info(void);            /* arch, compiler, flags, "src=<hash>" */
const char* deer_source_hash(void);           /* sha256 (hex) over csrc/ and include/ the library was built from; the binding
                                                  refuses a library whose hash differs from the sources beside it */
const char* deer_last_error(deer_ctx* ctx);   /* ctx may be NULL (creation errors) */

/* lifetime  -- replaces create_model_and_transforms + model.to(device) (factory.py:53, eval_calvin.py:565) */
int  deer_create(const deer_config* cfg, int device, deer_ctx** out);
void deer_destroy(deer_ctx* ctx);

/* weights -- replaces load_state_dict of the OpenFlamingo .pt + DeeR .pth (eval_calvin.py:541-578).
 * `name` is the reference's state-dict key (e.g. "perceiver.layers.0.0.to_q.weight",
 * "lang_encoder.transformer.blocks.3.decoder_layer.attn.Wqkv.weight", "extra_exit.rnn.layers.0.weight_ih_l0");
 * a leading "module." (DDP, train_utils.py:631-638) is ignored; unknown names return 1 (skipped, like strict=False).
 * `data` is a HOST pointer to contiguous fp32. */
int deer_load_weight(deer_ctx* ctx, const char* name, const float* data, const int64_t* shape, int ndim);
int deer_finalize_weights(deer_ctx* ctx);     /* checks completeness, packs fused operands, builds TMA descriptors */
int deer_num_required_weights(deer_ctx* ctx);
const char* deer_required_weight_name(deer_ctx* ctx, int i);

/* exit controller -- replaces ExitController._set_threshold_value (value_net.py:177-182).
 * n must equal the number of exits <= min(max_layer-1, E[-1]). */
int deer_set_thresholds(deer_ctx* ctx, const float* thresholds, int n);
int deer_num_exits(deer_ctx* ctx);            /* len(model.get_all_exit_idx()), flamingo_mpt.py:268-270 */
int deer_exit_layer(deer_ctx* ctx, int i);    /* i-th exit's layer index */
int deer_num_thresholds(deer_ctx* ctx);
int deer_ctrl_max_layer(deer_ctx* ctx);       /* value_net.py:173 */

/* episode state -- replaces ModelWrapper.reset -> clear_all_exit_memory (eval_utils.py:252-277,
 * flamingo_mpt.py:292-306): zero the LSTM (h,c) of the given slots (n < 0: all slots). HOST pointer. */
int deer_reset(deer_ctx* ctx, const int32_t* slots, int n);

/* one action-step for n concurrent episodes -- replaces ModelWrapper.step's model call +
 * post-processing (eval_utils.py:433-434, 454-477) for each episode independently.
 *   slots      [n]            which LSTM state each episode uses (distinct)
 *   rgb,gripper[n,3,H,W] f32  CLIP-normalised frames (what preprocess_image produces, data.py:898-902)
 *   ids        [n,S] int64    token ids (preprocess_text_calvin, data.py:905-919)
 *   mask       [n,S] uint8    attention mask (1 = token), right padded
 *   timesteps  [n] int32      rollout step of each episode (ExitController.set_timestep, value_net.py:274);
 *                             may be NULL (=0, always a stage boundary)
 *   static_exit               >=0: the reference's exit_id mode (flamingo_mpt.py:396-405), no probes; -1: dynamic
 * outputs (any may be NULL):
 *   pose_out   [n,6]  tanh head output              gripper_out [n] sigmoid head output
 *   action_out [n,7]  pose ++ (gripper>0.5 ? +1 : -1)   (eval_utils.py:458,464)
 *   exit_out   [n]    int32 exit layer index        delta_out [n, num_thresholds] (NaN where not evaluated)
 * deer_step: every array pointer is a DEVICE pointer, work is enqueued on `stream`
 * (a cudaStream_t), no host synchronisation unless cfg.host_control.
 * deer_step_host: every array pointer is a HOST pointer (pinned for full speed); copies in,
 * runs, copies out and synchronises before returning. */
int deer_step(deer_ctx* ctx, int n, const int32_t* slots, const float* rgb, const float* gripper,
              const int64_t* ids, const uint8_t* mask, int S, const int32_t* timesteps,
              int static_exit, uint32_t flags,
              float* pose_out, float* gripper_out, float* action_out, int32_t* exit_out, float* delta_out,
              void* stream);
int deer_step_host(deer_ctx* ctx, int n, const int32_t* slots, const float* rgb, const float* gripper,
                   const int64_t* ids, const uint8_t* mask, int S, const int32_t* timesteps,
                   int static_exit, uint32_t flags,
                   float* pose_out, float* gripper_out, float* action_out, int32_t* exit_out, float* delta_out);

/* measurement: with profiling on, every GEMM launch of the following steps is bracketed by CUDA events on the
 * launching stream, and the device-side active-episode count after every exit is recorded, so the exact
 * algorithmic FLOPs (2*M_actual*N*K) and the device time of the dominant kernel are known per stage.
 * stage index: 0 ViT, 1 Perceiver, 2 LLM (x-attn + MPT), 3 action head.  deer_get_profile synchronises. */
typedef struct deer_profile {
  double  gemm_ms[4];
  double  gemm_flops[4];
  int64_t gemm_launches[4];
  double  step_ms;            /* device time of the whole last step (first to last kernel) */
  int32_t n_probes;           /* exits evaluated in the last step */
  int32_t active_after[16];   /* still-active episodes after each probe */
  double  gemm_bytes[4];      /* algorithmic HBM bytes of those launches: A + W read once, C written once, fp32 residual read once */
} deer_profile;
int deer_set_profile(deer_ctx* ctx, int enable);
int deer_get_profile(deer_ctx* ctx, deer_profile* out);

/* Launch overlap (process-wide): 1 (default) launches the per-layer kernels with programmatic stream serialization, so a
 * kernel's prologue overlaps its predecessor's tail (the reference's eager PyTorch path serialises ~600 launches per step,
 * eval_utils.py:433); 0 = plain stream order.  Results are bit-identical either way (tested); returns the previous setting.
 * Environment DEER_PDL=0/1 sets the initial value. */
int deer_set_launch_overlap(int enable);

/* Row order of the streaming kernels that sit between two GEMMs (process-wide): 1 (default) = LayerNorm rows and ViT attention
 * items are walked from the END ("snake" order): they start on the rows the producing GEMM wrote last, which the 126 MB L2 still
 * holds, and finish on the rows the consuming GEMM reads first.  The kernels are row-independent, so results are bit-identical
 * either way (tested); returns the previous setting.  Environment DEER_SNAKE=0/1 sets the initial value. */
int deer_set_row_order(int snake);

/* One frame of the threshold-calibration pass in the reference's TRAINING layout -- replaces generate_action_values
 * (value_net.py:300-399): MPTFlamingo.forward(return_in_feat, only_extra_exit) over windows of frames (flamingo_mpt.py:476-492)
 * followed by ActionValueNet.forward(mode='generate') (value_net.py:134-160).  Call it for frames t = 0 .. window-2 of n
 * windows (slots reset before t = 0).  Every layer runs; the head probes layer 0 and every exit of the exit list from the
 * carried LSTM state; delta_out[i, k] = delta between ([0]+exit_list)[k] and [k+1] (DEVICE, [n, deer_num_exits()]); the
 * state carried to the next frame is the candidate computed at exit layer commit_layer[i] (DEVICE int32 [n]; the reference
 * draws it uniformly from the exit list per window and frame).  Deltas of frames window/2-1 .. window-2 form the table the
 * threshold solver consumes.  All pointers are DEVICE pointers. */
int deer_step_generate(deer_ctx* ctx, int n, const int32_t* slots, const float* rgb, const float* gripper, const int64_t* ids,
                       const uint8_t* mask, int S, const int32_t* commit_layer, float* delta_out, void* stream);

/* on-device image transform -- replaces preprocess_image -> open_clip image_transform for raw uint8 camera frames
 * (data.py:898-902; eval_utils.py:292-296,317-321): torchvision Resize(image_size, BICUBIC) on a PIL image (bit-exact
 * with Pillow's 8-bit resampler) -> CenterCrop -> /255 -> Normalize(CLIP mean/std).
 * frames: DEVICE uint8 [n,H,W,3]; out: DEVICE fp32 [n,3,image_size,image_size]. */
int deer_preprocess_u8(deer_ctx* ctx, const uint8_t* frames, int n, int H, int W, float* out, void* stream);
/* deer_step_host fed with RAW camera frames (HOST uint8 [n,Hs,Ws,3] static camera, [n,Hg,Wg,3] gripper camera, e.g.
 * CALVIN's 200x200 and 84x84): 8.5x fewer host->device bytes than normalised fp32 frames, transform on the GPU. */
int deer_step_host_u8(deer_ctx* ctx, int n, const int32_t* slots, const uint8_t* rgb, int Hs, int Ws, const uint8_t* gripper,
                      int Hg, int Wg, const int64_t* ids, const uint8_t* mask, int S, const int32_t* timesteps,
                      int static_exit, uint32_t flags,
                      float* pose_out, float* gripper_out, float* action_out, int32_t* exit_out, float* delta_out);

/* introspection for parity tests (the reference exposes .hidden_states / module attributes) */
int deer_get_state(deer_ctx* ctx, int slot, float* h_out, float* c_out);   /* HOST [layers, hidden] each */
/* replaces assigning `model.extra_exit.hidden_state = (h, c)` (action_head.py:474,554-558): overwrite one slot's carried
 * LSTM state.  HOST pointers [layers, hidden]; ordered after the last step like deer_reset. */
int deer_set_state(deer_ctx* ctx, int slot, const float* h, const float* c);
/* `.logits[2]` of the reference's output (return_feature=True: the max-pooled token feature the exit head consumed,
 * action_head.py:519-520,608-609; flamingo_mpt.py:458-460) for the n episodes of the LAST step, row i = episode i of that
 * call, taken at each episode's own exit layer.  out: [n, d_model] fp32, DEVICE (copied on `stream`) or HOST
 * (out_is_host = 1; synchronises). */
int deer_get_features(deer_ctx* ctx, int n, float* out, int out_is_host, void* stream);
/* name: "vit" [2n, patches, width] (rgb images then gripper images), "media" [n, 2*latents, width],
 * "hidden" [n, S, d] (layer = LLM layer index; rows of episodes that exited earlier keep their last value),
 * "pooled" [n, d].  Requires cfg.debug_taps.  out is a HOST buffer of `capacity` floats; shape_out[4]. */
int deer_get_tap(deer_ctx* ctx, const char* name, int layer, float* out, int64_t capacity, int64_t* shape_out);
int deer_get_stats(deer_ctx* ctx, deer_stats* out);

/* single-operator entry points (DEVICE pointers) so that kernels can be parity-tested in isolation.
 * dtype: 0 = bf16, 1 = fp32 for A/W (and C unless out_f32).   C[M,N] = epi(A[M,K] @ W[N,K]^T).
 * epilogue: v = acc (+bias[n]); act 0 none / 1 GELU(erf) / 2 QuickGELU / 3 ReLU; v *= *scale (if scale);
 * v += resid[m,n] (fp32, ldc) (if resid).  m_dev (optional device int) clamps M to *m_dev * m_mul; rows at or beyond that count are
 * never written for fp32 / residual outputs, while a bf16 output without residual leaves through whole 32-row TMA tiles, so its rows
 * between the device-side count and M are scratch (they may be overwritten; rows >= M never are). */
int deer_op_gemm(int impl, int dtype, const void* A, int lda, const void* W, int ldw, void* C, int ldc,
                 int M, int N, int K, const float* bias, int act, const float* scale, const float* resid,
                 int out_f32, const int32_t* m_dev, int m_mul, void* stream);
/* softmax(scale * Q K^T) V, no mask (impl 0 = tensor-core path: the tcgen05 kernel when Sk <= 272, else the mma.sync flash
 * kernel; 1 = SIMT reference; 3 = force the mma.sync flash kernel; dtype as above, impl 0/3 require bf16 and head_dim 64).  Q rows: (b*Sq+i)*ldq + h*64, K/V rows: (b*Sk+j)*ldk + h*64. */
int deer_op_attention(int impl, int dtype, const void* Q, int ldq, const void* K, const void* V, int ldk,
                      void* O, int ldo, int B, int H, int Sq, int Sk, int head_dim, float scale, void* stream);

/* the LLM-side attentions with their masks (gated x-attn: helpers.py:205-229; MPT self-attention: mosaic_gpt_3b.py:158-219):
 * bias = alibi_slopes[h] * (j - (Sk-1)) (or none), causal (key j visible to query i iff j <= i + Sk - Sq), kv_len[b] valid keys,
 * qmode[b,i] in {0: zero output, 1: attend, 2: uniform}, n_dev = device-side episode count (<= B) or NULL.  All optional
 * pointers are DEVICE pointers or NULL.  impl 1 = SIMT reference, 4 = mma.sync kernel (bf16, Sq<=32, Sk<=128, head_dim 64/128),
 * 2 = what the engine picks. */
int deer_op_attention_masked(int impl, int dtype, const void* Q, int ldq, const void* K, const void* V, int ldk, void* O, int ldo,
                             int B, int H, int Sq, int Sk, int head_dim, float scale, int causal, const float* alibi_slopes,
                             const int32_t* kv_len, const uint8_t* qmode, const int32_t* n_dev, void* stream);

/* bring-up / regression probe of the tensor-core attention's second GEMM: O[128,64] = P[128,K] @ V[K,64] with P staged in
 * tensor memory (tcgen05.st) as the A operand and V read in its natural [key][d] layout as an MN-major B operand. */
int deer_op_pv_probe(const void* P, const void* V, float* O, int K, int lbo_bytes, int sbo_bytes, int kstep_bytes, void* stream);

/* tuning aid: per-CTA cycle counters of the attention kernel's softmax warp (wait-S, pass 1, pass 2, wait-O, epilogue). */
int deer_attn_debug(int enable, long long* host_out, int n);
/* tuning aid: %globaltimer (ns) of CTA 0 at the phase boundaries of the last fused exit-head probe (head_fused.cuh): [0] entry,
 * per grid barrier k = 1 .. lstm_layers + 3 the pair ([2k-1] arrival, [2k] release), then the exit. */
int deer_head_debug(int enable, long long* host_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* DEER_ENGINE_H */
