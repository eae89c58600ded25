/* pnpinv.h -- C ABI of libpnpinv.so: the B200-native (sm_100a) hot path of cure-lab/PnPInversion.
 *
 * The reference has no FFI: its "plugin API" is three duck-typed Python seams.  This header
 * is what a replacement for seam B (the `model` handle the loops call) and seam C (the attention controller) binds
 * to.  Every entry point cites the reference interface it replaces.  Conventions:
 *   - all functions return 0 on success, <0 on error; the message is available from pnp_last_error();
 *   - nothing throws across the boundary; a handle is not thread-safe; one handle per GPU/process;
 *   - every launch goes to the caller-supplied cudaStream_t (passed as void*; 0 = legacy default stream);
 *   - pointers named *_dev are device pointers borrowed for the duration of the call (until the stream reaches it),
 *     pointers named *_host are host pointers read before the call returns;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef PNPINV_H_
#define PNPINV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNP_MAX_BATCH 32   /* UNet batch rows (CFG pair included) */
#define PNP_MAX_SLOTS 8    /* edited (target) rows with their own 77-entry tables */
#define PNP_TOKENS 77      /* MAX_NUM_WORDS, models/p2p/attention_control.py:8 */
#define PNP_LATENT_ELEMS 16384 /* 4 x 64 x 64 */

typedef struct pnp_engine pnp_engine;

/* ---- lifecycle ---------------------------------------------------------------------------------------------- */
/* replaces StableDiffusionPipeline.from_pretrained(...).to(device) for the UNet part, models/p2p_editor.py:23-24 */
int pnp_create(int device_ordinal, int max_batch, pnp_engine** out);
/* a second handle on the same GPU that SHARES the parent's read-only parameter buffers and time-embedding table (1.72 GB)
 * and owns only its activation arenas, CUDA graphs, controller state and stream: several images in flight per GPU
 * without replicating the weights.  The parent must be finalized, must have seen pnp_set_timesteps, and must outlive
 * the clone. */
int pnp_clone(pnp_engine* parent, int max_batch, pnp_engine** out);
void pnp_destroy(pnp_engine* h);
const char* pnp_last_error(void);
const char* pnp_version(void);

/* ---- parameters: the 686 tensors of the SD-1.x UNet state_dict, names verbatim --------- */
int pnp_unet_param_count(void);
/* name_out: >=128 bytes; shape_out: 4 ints (unused dims = 0) */
int pnp_unet_param_spec(int index, char* name_out, int* ndim_out, int* shape_out);
/* data_host: fp16 values in PyTorch layout (conv OIHW, linear (out,in)); copied before returning */
int pnp_load_param(pnp_engine* h, const char* name, const uint16_t* data_host, int64_t numel);
/* repacks everything once into the kernel layouts (conv -> (Cout, 9*Cin [+Cin_shortcut]) K-major, fused QKV, GEGLU
 * tile interleave) and uploads; all 686 tensors must have been loaded */
int pnp_finalize_params(pnp_engine* h);

/* ---- per-schedule / per-prompt precomputation --------------------------------------------------------------- */
/* replaces Timesteps + TimestepEmbedding + every ResnetBlock2D.time_emb_proj evaluated per UNet call
 * (my_diffusers/models/embeddings.py:21-80, resnet.py:348-350): computed once for the n distinct timestep values */
int pnp_set_timesteps(pnp_engine* h, const int64_t* timesteps_host, int n, void* stream);
/* replaces attn2.to_k / attn2.to_v on encoder_hidden_states in all 16 cross-attention layers
 * (my_diffusers/models/attention.py:255-257): context is step-invariant, so K/V are computed once.
 * ctx_dev: [batch, 77, 768] fp32 */
int pnp_set_context(pnp_engine* h, const float* ctx_dev, int batch, void* stream);

/* ---- seam C: the attention controller lowered to a descriptor ------------------------------------------------ */
/* replaces controller(attn, is_cross, place_in_unet) (models/p2p/attention_control.py:43-45,178-190,269-282) and
 * editor(q,k,v,sim,attn,...) (models/masactrl/masactrl_utils.py:119-121).  Rows are UNet batch rows. */
typedef struct pnp_attn_ctrl {
  /* self-attention: in transformer blocks [self_layer_lo, self_layer_hi) with <= self_max_tokens tokens, batch row r
   * takes Q from row self_q_row[r], K from self_k_row[r], V from self_v_row[r] (identity = r). */
  int32_t self_layer_lo, self_layer_hi, self_max_tokens;
  int32_t self_q_row[PNP_MAX_BATCH];
  int32_t self_k_row[PNP_MAX_BATCH];
  int32_t self_v_row[PNP_MAX_BATCH];
  /* cross-attention: row r with cross_base_row[r] >= 0 is a target row whose probabilities become
   *   ((P_src[:, mapper] * alphas + P_r * (1 - alphas)) * equalizer) * cross_alpha + (1 - cross_alpha) * P_r
   * with the tables of slot cross_slot[r]  (AttentionRefine / AttentionReweight / time gate). */
  int32_t cross_base_row[PNP_MAX_BATCH];
  int32_t cross_slot[PNP_MAX_BATCH];
  int32_t mapper[PNP_MAX_SLOTS][PNP_TOKENS];
  float alphas[PNP_MAX_SLOTS][PNP_TOKENS];
  float equalizer[PNP_MAX_SLOTS][PNP_TOKENS];
  float cross_alpha[PNP_MAX_SLOTS][PNP_TOKENS];
  /* AttentionReplace with unequal token spans (models/p2p/seq_aligner.py:152-185, applied by the einsum of
   * attention_control.py:303-304): column w of the 77x77 mapper is `map_weight` on the map_count[w] consecutive source
   * tokens starting at mapper[w], so P_src[:, mapper] above generalises to
   *   map_weight[w] * sum_{k < map_count[w]} P_src[:, mapper[w] + k]          (identity: count 1, weight 1). */
  int32_t map_count[PNP_MAX_SLOTS][PNP_TOKENS];
  float map_weight[PNP_MAX_SLOTS][PNP_TOKENS];
  /* Plug-and-Play feature injection (run_editing_pnp.py:244-294): the conv2 output of up_blocks.1.resnets.1 of row r is
   * computed from the hidden state of row conv_src_row[r] (the residual / shortcut term stays row r's own); identity = r */
  int32_t conv_src_row[PNP_MAX_BATCH];
  /* AttentionStore for LocalBlend: rows with store_slot[r] >= 0 accumulate their (post-injection) 16x16 cross maps of
   * the five layers down_cross[2:4] + up_cross[:3] into slot store_slot[r]. */
  int32_t store_slot[PNP_MAX_BATCH];
} pnp_attn_ctrl;
void pnp_attn_ctrl_init(pnp_attn_ctrl* c); /* identity / disabled everywhere */

/* ---- seam B: model.unet(latents, t, encoder_hidden_states=ctx)["sample"] ------------------------------------- */
/* replaces UNet2DConditionModel.forward (my_diffusers/models/unet_2d_condition.py:189-273; called at
 * models/p2p/inversion.py:273, p2p_guidance_forward.py:109).  x_dev/eps_out_dev: [batch,4,64,64] fp32 NCHW;
 * t_index: index into the list given to pnp_set_timesteps; context from pnp_set_context (same batch);
 * ctrl_host may be NULL (= no controller). */
int pnp_unet_forward(pnp_engine* h, const float* x_dev, int batch, int t_index, const pnp_attn_ctrl* ctrl_host,
                     float* eps_out_dev, void* stream);

/* ---- the fused step epilogue: CFG + DDIM (inverse) step + the "3 lines" -------------------------------------- */
/* replaces next_step / prev_step / DDIMSchedulerDev.step, the CFG combine, `loss = latent_prev - rec` and
 * `latents[:1] + noise_loss[:1]` (models/p2p/inversion.py:247-270,383-389; scheduler_dev.py:40-51,91-94;
 * p2p_guidance_forward.py:111-114).  All coefficient scalars are computed by the caller from the alphas_cumprod
 * table exactly as the reference does (integer timestep arithmetic stays on the host, bit-exact). */
typedef struct pnp_step_args {
  const float* x_dev;      /* [n,16384] current latents */
  const float* eps_u_dev;  /* [n,16384] unconditional prediction, or NULL (no guidance: eps = eps_c) */
  const float* eps_c_dev;  /* [n,16384] */
  float* x_out_dev;        /* [n,16384] (may alias x_dev) */
  int32_t n;
  float guidance;
  float sqrt_a_from, sqrt_1m_a_from, sqrt_a_to, sqrt_1m_a_to;
  const float* target_dev; /* offset mode: [target_rows,16384], or NULL */
  int32_t target_rows;
  float* loss_out_dev;     /* offset mode: [n,16384] */
  float loss_scale;        /* offset mode: loss = (target - x_new) * loss_scale; 1 = DirectInversion.offset_calculate,
                              `scale` of offset_calculate_not_full (inversion.py:478-492), 0 on the skipped steps of
                              offset_calculate_skip_step (:501-519) */
  const float* noise_loss_dev; /* rectification: [n,16384], or NULL */
  uint32_t add_mask;       /* bit r: add noise_loss row r */
} pnp_step_args;
int pnp_step_epilogue(pnp_engine* h, const pnp_step_args* a, void* stream);

/* replaces LocalBlend.__call__ (models/p2p/attention_control.py:97-121) on the latents [2,4,64,64] (row 0 source,
 * row 1 target) using the maps accumulated through pnp_attn_ctrl.store_slot {0,1}. words/alpha: the non-zero
 * entries of alpha_layers per prompt (<= 8 each). mask_out_dev: optional [2,4096] floats. */
int pnp_local_blend(pnp_engine* h, float* x_dev, const int32_t* nwords2_host, const int32_t* words2x8_host,
                    const float* alpha2x8_host, float threshold, float* mask_out_dev, void* stream);
/* the same for several (source, target) latent pairs of one batch in ONE launch, with the `substruct_words` branch
 * (attention_control.py:116-118): mask = pooled-mask(words, th_pool) & ~unpooled-mask(sub_words, th_sub). */
typedef struct pnp_blend_desc {
  int32_t src_row, tgt_row;   /* latent rows of the pair inside x_dev */
  int32_t src_slot, tgt_slot; /* store slots the two branches accumulated into (pnp_attn_ctrl.store_slot) */
  int32_t nwords[2];          /* [source prompt, target prompt] */
  int32_t words[2][8];
  float alpha[2][8];
  int32_t nsub[2];            /* substruct words (0 = none) */
  int32_t sub_words[2][8];
  float sub_alpha[2][8];
  float th_pool, th_sub;      /* LocalBlend.th[0], th[1] */
} pnp_blend_desc;
#define PNP_MAX_BLEND 8
int pnp_local_blend_batch(pnp_engine* h, float* x_dev, int n_rows, const pnp_blend_desc* descs_host, int n_desc,
                          float* mask_out_dev /* optional [n_desc][2][4096] */, void* stream);

/* ---- whole step loops behind the boundary (pnp_invert / pnp_offset / pnp_edit) ---------- */
/* One call enqueues `n_steps` x { UNet forward ; fused step epilogue [; LocalBlend] } on the stream: no per-step host
 * tensor op, no intermediate copy (the epilogue reads the UNet's output buffer in place).  Latent rows are
 * PROMPT-MAJOR for L images with P prompts each: row = p * L + image; the UNet batch of the guided modes is
 * [unconditional rows | conditional rows] like `torch.cat([latents] * 2)` (p2p_guidance_forward.py:108).
 *   PNP_LOOP_INVERT  replaces DirectInversion.ddim_loop (models/p2p/inversion.py:308-319; with guidance != 0 the
 *                    CFG variant ddim_with_guidance_scale_loop :349-363): step i runs the UNet at t_host[i] and the
 *                    inverse step with coef_host[i]; traj_dev receives the n_steps+1 latents (x_stars).
 *   PNP_LOOP_OFFSET  replaces DirectInversion.offset_calculate (:375-391): target of step i = traj_dev[n_steps-i-1]
 *                    (row r uses image r % images), loss_dev[i] = (target - rec) * loss_scale_host[i], x = rec + loss.
 *   PNP_LOOP_FORWARD replaces direct_inversion_p2p_guidance_forward (models/p2p/p2p_guidance_forward.py:135-173) and,
 *                    with loss_dev == NULL, p2p_guidance_forward (:21-62): CFG + scheduler.step + rectification of the
 *                    rows in add_mask with loss_dev[i], controller descriptor ctrl_host[i], LocalBlend after step i
 *                    once i + 1 > blend_start. */
#define PNP_LOOP_INVERT 0
#define PNP_LOOP_OFFSET 1
#define PNP_LOOP_FORWARD 2
typedef struct pnp_loop_args {
  int32_t mode, n_steps;
  int32_t rows;                /* latent rows n (UNet batch n for INVERT without guidance, else 2n) */
  int32_t images;              /* L */
  const int32_t* t_host;       /* [n_steps] timestep of each step's UNet call */
  const float* coef_host;      /* [n_steps][4]: sqrt_a_from, sqrt_1m_a_from, sqrt_a_to, sqrt_1m_a_to */
  float guidance;
  const float* ctx_dev;        /* [UNet batch,77,768] fp32 */
  float* x_dev;                /* [rows,16384] start latents in, final latents out */
  float* traj_dev;             /* INVERT: out [n_steps+1][rows][16384]; OFFSET: in [n_steps+1][images][16384] */
  float* loss_dev;             /* OFFSET: out [n_steps][rows][16384]; FORWARD: in, or NULL (no rectification) */
  const float* loss_scale_host;/* OFFSET: [n_steps], or NULL (= 1) */
  uint32_t add_mask;           /* FORWARD: rows that receive loss_dev */
  const pnp_attn_ctrl* ctrl_host; /* FORWARD: [n_steps] descriptors, or NULL */
  const pnp_blend_desc* blend_host; /* FORWARD: [n_blend] or NULL */
  int32_t n_blend, blend_start;
} pnp_loop_args;
int pnp_run_loop(pnp_engine* h, const pnp_loop_args* a, void* stream);

/* replaces the EDICT mixing layers on the coupled latent pair, in place (models/edict/edict_functions.py:854-859 when
 * reverse != 0, :931-936 otherwise); x_dev, y_dev: [n_rows,16384] */
int pnp_edict_mix(pnp_engine* h, float* x_dev, float* y_dev, int n_rows, float mix_weight, int reverse, void* stream);
int pnp_store_reset(pnp_engine* h, void* stream);          /* AttentionStore.reset() */
/* debug/inspection: copy the accumulated maps [5][2*PNP_MAX_SLOTS... see DESIGN.md] */
int pnp_store_read(pnp_engine* h, float* out_dev, int64_t max_floats, void* stream);

/* inspection: the activations the last forward of this batch size left in skip tensor `which` (0..11: conv_in output,
 * then every down-path block output; fp16 NHWC [batch, H, W, C]) - used by tools/diag_layers.py */
int pnp_debug_read(pnp_engine* h, int batch, int which, uint16_t* out_dev, int64_t max_elems, int64_t* n_out, void* stream);

/* sizeof() of the boundary structs as this library was compiled (0 pnp_attn_ctrl, 1 pnp_step_args, 2 pnp_blend_desc,
 * 3 pnp_loop_args): lets a binding verify its mirror of the layouts at load time. */
int pnp_struct_size(int which);

/* ---- instrumentation ----------------------------------------------------------------------------------------- */
/* runs one UNet forward eagerly with a CUDA event pair around every op of the plan (on the engine's stream);
 * kind: 0 tcgen05 gemm/conv, 1 groupnorm, 2 layernorm, 3 self-attention, 4 cross-attention, 5 other;
 * flops: algorithmic 2*MAC of the op.  Used by bench.py for the roofline of the dominant kernel. */
int pnp_unet_profile(pnp_engine* h, int batch, int t_index, int reps, float* ms_out, int32_t* kind_out,
                     double* flops_out, int max_ops, int* n_out); /* each op launched `reps` times back to back */
/* compulsory HBM bytes of all GEMM / implicit-conv launches of one UNet call at this batch (activations, weights and
 * residual read once, output written once) and their number: the algorithmic figure ncu's DRAM traffic is held against */
int pnp_unet_gemm_bytes(pnp_engine* h, int batch, double* bytes_out, int* launches_out);
int pnp_kernel_launches(pnp_engine* h, int64_t* out); /* kernels launched by this handle so far (graph nodes count) */
int pnp_set_use_graph(pnp_engine* h, int enable);     /* capture each UNet forward into a CUDA graph (default on) */

/* ---- the VAE around the loop ------------------------------------------------------- */
/* replaces AutoencoderKL.encode / .decode as utils/utils.py:58-80 calls them (image2latent: encode(img).latent_dist.mean
 * * 0.18215; latent2image: decode(z / 0.18215)); arithmetic spec my_diffusers/models/vae.py:54-210,480-557.  Parameters
 * carry the names of a diffusers SD-1.x `vae/` state dict (248 tensors, fp16 values in PyTorch layouts).  The 0.18215
 * scaling and the uint8 conversion stay with the caller, exactly where the reference has them. */
typedef struct pnp_vae pnp_vae;
int pnp_vae_create(int device_ordinal, pnp_vae** out);
void pnp_vae_destroy(pnp_vae* h);
int pnp_vae_load_param(pnp_vae* h, const char* name, const uint16_t* data_host, int64_t numel);
int pnp_vae_finalize(pnp_vae* h);
/* image_dev: [batch,3,H,W] fp32 in [-1,1] (H, W in {128,256,512}); moments_out_dev: [batch,8,H/8,W/8] fp32 = posterior
 * mean (channels 0..3) and log-variance (4..7) after quant_conv */
int pnp_vae_encode(pnp_vae* h, const float* image_dev, int batch, int img_h, int img_w, float* moments_out_dev,
                   void* stream);
/* z_dev: [batch,4,h,w] fp32 (already divided by 0.18215); image_out_dev: [batch,3,8h,8w] fp32 */
int pnp_vae_decode(pnp_vae* h, const float* z_dev, int batch, int lat_h, int lat_w, float* image_out_dev, void* stream);
int pnp_vae_kernel_launches(pnp_vae* h, int64_t* out);

/* ---- CLIP text encoder (`model.text_encoder(input_ids)[0]`: models/p2p/inversion.py:42,50,296,304,
 * models/p2p/p2p_guidance_forward.py:43,49,86,92, models/edict/edict_functions.py:818-838) ------------------------
 * The reference takes it from `transformers` (CLIPTextModel of the SD-1.x checkpoint: 768 hidden, 12 pre-LayerNorm
 * blocks of 12 heads, quick_gelu MLP of 3072, 77 positions, causal mask, final LayerNorm).  Parameters are loaded under
 * the names of that model's state_dict (`text_model.embeddings.token_embedding.weight`, `text_model.encoder.layers.<i>.
 * self_attn.q_proj.weight`, ...) as fp16 bits; the layer count and the vocabulary size are taken from what was loaded. */
typedef struct pnp_clip pnp_clip;
int pnp_clip_create(int device_ordinal, pnp_clip** out);
void pnp_clip_destroy(pnp_clip* h);
int pnp_clip_load_param(pnp_clip* h, const char* name, const uint16_t* data_host, int64_t numel);
int pnp_clip_finalize(pnp_clip* h);
int pnp_clip_vocab_size(pnp_clip* h, int* vocab_out, int* layers_out);
/* input_ids_host: [batch,77] int32 token ids in HOST memory (validated against the vocabulary: an id outside it is an
 * error, as the embedding lookup of the reference raises); out_dev: [batch,77,768] fp32 last_hidden_state */
int pnp_clip_encode(pnp_clip* h, const int32_t* input_ids_host, int batch, float* out_dev, void* stream);
int pnp_clip_kernel_launches(pnp_clip* h, int64_t* out);

/* ---- stand-alone kernel entry points (used by tests/ and bench.py to measure single kernels) ------------------ */
/* D[M,N] = A[M,K].W[N,K]^T (+bias)(+residual) ; mode 0 plain, 1 GEGLU (N = 2*out columns, weights pre-interleaved by
 * pnp_test_pack_geglu) ; conv3x3: A is NHWC [B,H,W,C], W packed (N, 9*C) tap-major */
int pnp_test_gemm(const uint16_t* a_dev, int M, int K, int lda, const uint16_t* w_dev, int N, const float* bias_dev,
                  const uint16_t* residual_dev, uint16_t* out_dev, int ldc, int geglu, int bn, int split,
                  void* stream); /* split: 0 auto, 1 off, n>1 force n K-splits */
/* D[M,N] = [A0 | A1][M, K0+K1] . W[N, K0+K1]^T: two A sources along K (how a hi/lo split of the activation operand runs on
 * the unchanged kernel: A0 = fp16(A), A1 = fp16(A - A0), W = [W | W]); reps > 1 launches it that many times and returns
 * the mean device time per launch in *ms_out (CUDA events) */
int pnp_test_gemm2(const uint16_t* a0_dev, int K0, const uint16_t* a1_dev, int K1, int M, const uint16_t* w_dev, int N,
                   uint16_t* out_dev, int reps, float* ms_out, void* stream);
int pnp_test_conv3x3(const uint16_t* x_dev, int B, int H, int W, int C, const uint16_t* w_dev, int N,
                     const uint16_t* sc0_dev, int sc0_C, const uint16_t* sc1_dev, int sc1_C, const float* bias_dev,
                     const uint16_t* residual_dev, uint16_t* out_dev, int bn, int split, void* stream);
int pnp_test_groupnorm(const uint16_t* x0_dev, int C0, const uint16_t* x1_dev, int C1, int B, int HW,
                       const float* gamma_dev, const float* beta_dev, float eps, int silu, uint16_t* out_dev,
                       void* stream);
/* which GroupNorm path groupnorm_launch takes for this shape on the current device: 0 = statistics + apply (two kernels),
 * 1 = single-launch cluster kernel, 2 = register-resident kernel (one CTA per image x group chunk) */
int pnp_test_groupnorm_path(int C, int B, int HW);
int pnp_test_layernorm(const uint16_t* x_dev, int rows, int C, const float* gamma_dev, const float* beta_dev,
                       float eps, uint16_t* out_dev, void* stream);
int pnp_test_self_attention(const uint16_t* qkv_dev, int B, int H, int N, int d, const int32_t* q_row_dev,
                            const int32_t* k_row_dev, const int32_t* v_row_dev, uint16_t* out_dev, void* stream);
/* tcgen05 path (8 heads of dim 40, N % 128 == 0); qkv_dev: [B,N,960] */
int pnp_test_self_attention_tc(const uint16_t* qkv_dev, int B, int N, const int32_t* q_row_dev, const int32_t* k_row_dev,
                               const int32_t* v_row_dev, uint16_t* out_dev, void* stream);
int pnp_test_cross_attention(const uint16_t* q_dev, const uint16_t* kv_dev, int B, int H, int N, int d, int nk,
                             const pnp_attn_ctrl* ctrl_host, float* store_dev, uint16_t* out_dev, void* stream);
/* tcgen05.mma instruction-cost probe (csrc/probe.cu, tools/mma_probe.py): `n` back-to-back MMAs of shape M x N x 16
 * (cta_group 1: M = 64 / 128; cta_group 2: M = 256 on a cluster of two SMs), A operand from tensor memory or shared
 * memory, round-robin over `nacc` accumulators, issued in elected blocks of `group` (1..8) instructions, `commit` = a
 * tcgen05.commit after every block; cycles_out_host[0] = cycles to issue them, [1] = cycles until the final commit arrives */
int pnp_test_mma_probe(int cta_group, int M, int N, int a_from_tmem, int n, int nacc, int group, int commit,
                       int64_t* cycles_out_host);
int pnp_test_upsample2x(const uint16_t* x_dev, int B, int H, int W, int C, uint16_t* out_dev, void* stream);
int pnp_test_im2col_s2(const uint16_t* x_dev, int B, int H, int W, int C, uint16_t* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PNPINV_H_ */
