// Parameter blocks of the attention kernels (attention.cu) and of the fused step epilogue (epilogue.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace pnp {

// q/k/v point into one fused [B, N, ld] buffer (ld = 3C): q at +0, k at +C, v at +2C; head h at column h*d.
// *_row (device, length B, may be null = identity) implement the controllers as batch-row indirection:
//   P2P self-replace : q_row[tgt]=src, k_row[tgt]=src            (attention_control.py:258-263)
//   MasaCtrl         : k_row[r]=src(r), v_row[r]=src(r)          (masactrl.py:56-72)
struct SelfAttnParams {
  const __half* q;
  const __half* k;
  const __half* v;
  int ld;
  __half* o;
  int ldo;
  int B, H, N, d;
  float scale;
  const int* q_row;
  const int* k_row;
  const int* v_row;
};

// q: [B, N, ldq]; kv: [B, nk, ldkv] with K at +0 and V at +C (C = H*d), precomputed once per prompt.
// Per batch row r: base_row[r] >= 0 marks an edited (target) row whose source row is base_row[r]; edit_slot[r]
// selects its 77-entry tables (mapper/alphas/equalizer/cross_alpha).  store (may be null): accumulate the final
// probabilities of rows with store_slot[r] >= 0 into store[slot][h][q][77] (the maps LocalBlend consumes).
struct CrossAttnParams {
  const __half* q;
  int ldq;
  const __half* kv;
  int ldkv;
  __half* o;
  int ldo;
  int B, H, N, d, nk;
  float scale;
  const int* base_row;
  const int* edit_slot;
  const int* mapper;         // [slots][77] int32
  const float* alphas;       // [slots][77]
  const float* equalizer;    // [slots][77]
  const float* cross_alpha;  // [slots][77]  (row cur_step of cross_replace_alpha)
  const int* map_count;      // [slots][77] source tokens summed per target token (null = 1), see pnp_attn_ctrl
  const float* map_weight;   // [slots][77] weight of that sum (null = 1)
  float* store;
  const int* store_slot;
  int tiles_per_cta = 1;  // consecutive 64-query tiles one CTA walks (set by the launcher)
};

int self_attention_launch(const SelfAttnParams& p, cudaStream_t s);

// tcgen05 path for the 4096-token layers (8 heads of dim 40), attention_tc.cu
struct alignas(64) SelfAttnTcParams {
  CUtensorMap map_qk;  // fused QKV activation viewed as [B*N][3][8][40]
  CUtensorMap map_vt;  // V^T scratch [B][8][41][N] (row 40 = ones)
  CUtensorMap map_k64;  // same view as map_qk with a 64-row box (cluster-of-2 multicast: each CTA loads half a K tile)
  CUtensorMap map_vt32;  // V^T with a 32-row box (pair mode: each CTA stages half of the 64 columns of O)
  int cluster;          // 1 = one CTA per query tile, 2 = cluster of two with TMA multicast, 3 = pair (tcgen05.mma.cta_group::2)
  int poly;             // packed exponentials per 8 evaluated on the FMA pipe instead of the MUFU (0, 2, 3, 4)
  int sched;            // MMA issue order: 0 = fixed (S(j+2), then P(j) V(j)), 1 = event-driven (whichever has its inputs, S first)
  int roles_hi;         // 1 = TMA / MMA / allocator roles on the highest warp ids (issue priority), softmax warps on 0..15
  const __half* q_src;  // Q part of the fused activation (rows are copied into TMEM by the kernel)
  const __half* v_src;  // V part of the fused activation (transpose source)
  int ld;
  __half* vt;
  __half* o;
  int ldo;
  int B, N;
  float sl2;  // scale * log2(e)
  const int* q_row;
  const int* k_row;
  const int* v_row;
  volatile unsigned int* dbg;
  // optional cycle counters [CTA][16] (null in production): MMA warp total / wait k_full / s_empty / p_full / v_full,
  // producer total / wait k_empty / v_empty, softmax warp 4 (group 0) and 12 (group 1): total / wait s_full / p_empty
  long long* prof;
};
size_t self_attention_tc_vt_elems(int B, int N);
int self_attention_tc_init_vt(__half* vt, int B, int N, cudaStream_t s);
int self_attention_tc_plan(SelfAttnTcParams* p, const __half* qkv, int ld, __half* vt, __half* o, int ldo, int B, int N,
                           const int* q_row, const int* k_row, const int* v_row);
int self_attention_tc_launch(const SelfAttnTcParams& p, cudaStream_t s);
int cross_attention_launch(const CrossAttnParams& p, cudaStream_t s);

// ------------------------------------------------------------------ fused step epilogue (epilogue.cu)
// One launch per diffusion step over `n` latent rows of 4*64*64 fp32:
//   eps   = eps_u + g * (eps_c - eps_u)                       p2p_guidance_forward.py:111 ; inversion.py:282
//   x0    = (x - sqrt(1-a_from) * eps) / sqrt(a_from)         inversion.py:252-253 / 266-267 ; scheduler_dev.py:46,51
//   x_new = sqrt(a_to) * x0 + sqrt(1-a_to) * eps              inversion.py:254-255 / 268-269 ; scheduler_dev.py:91-94
//   OFFSET  : loss = target - x_new ; x_new = x_new + loss     inversion.py:386-389
//   RECTIFY : x_new += noise_loss (rows flagged in add_mask)   p2p_guidance_forward.py:113-114
struct StepParams {
  const float* x;      // [n, 16384]
  const float* eps_u;  // [n, 16384] (null: no CFG, eps = eps_c)
  const float* eps_c;  // [n, 16384]
  float* x_out;        // [n, 16384]
  int n;
  float guidance;
  // fp32 scalars computed on the host exactly the way the reference does (alphas_cumprod table in fp32)
  float sqrt_a_from, sqrt_1m_a_from;  // x0 = (x - sqrt_1m_a_from * eps) / sqrt_a_from
  float sqrt_a_to, sqrt_1m_a_to;      // x_new = sqrt_a_to * x0 + sqrt_1m_a_to * eps
  const float* target;      // OFFSET: [target_rows, 16384] latent the branch must land on (row r uses r % target_rows)
  int target_rows;
  float* loss_out;          // OFFSET: [n, 16384]
  float loss_scale;         // OFFSET: loss = (target - x_new) * loss_scale (1 = the default method)
  const float* noise_loss;  // RECTIFY: [n, 16384]
  unsigned add_mask;        // RECTIFY: bit r set -> add noise_loss row r to x_new row r
};
int step_epilogue_launch(const StepParams& p, cudaStream_t s);
int edict_mix_launch(float* x, float* y, int n_elems, float w, bool reverse, cudaStream_t s);

// LocalBlend (attention_control.py:97-121), one CTA per (source, target) latent pair
struct LocalBlendItem {
  int src_row, tgt_row, src_slot, tgt_slot;
  int nwords[2];  // words with alpha_layers != 0 per prompt
  int words[2][8];
  float alpha[2][8];
  int nsub[2];  // substruct words (attention_control.py:116-118)
  int sub_words[2][8];
  float sub_alpha[2][8];
  float th_pool, th_sub;
};
struct LocalBlendParams {
  const float* store;  // [5 layers][slots][8 heads][256 queries][77] running sum over steps
  long long layer_stride, slot_stride;  // floats
  float* x;         // [rows, 4, 64, 64] in/out
  float* mask_out;  // optional [n_items][2][64*64] inspection output (may be null)
  int n_items;
  LocalBlendItem items[8];
};
int local_blend_launch(const LocalBlendParams& p, cudaStream_t s);

}  // namespace pnp
