// Internal C++ declarations shared by the kernels and the engine (not part of the C ABI; see include/pnpinv.h).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace pnp {

// ------------------------------------------------------------------ errors
void set_last_error(const std::string& msg);
const char* get_last_error();
#define PNP_CUDA(call)                                                                                          \
  do {                                                                                                          \
    cudaError_t e__ = (call);                                                                                   \
    if (e__ != cudaSuccess) {                                                                                   \
      ::pnp::set_last_error(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                            std::to_string(__LINE__));                                                          \
      return -1;                                                                                                \
    }                                                                                                           \
  } while (0)
#define PNP_CHECK(cond, msg)                                                                    \
  do {                                                                                          \
    if (!(cond)) {                                                                              \
      ::pnp::set_last_error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" + std::to_string(__LINE__)); \
      return -2;                                                                                \
    }                                                                                           \
  } while (0)

// host-mapped debug words written by a kernel that detects a stuck pipeline (see mbar_wait)
volatile unsigned int* debug_words_device();  // device-visible alias
const unsigned int* debug_words_host();

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// Every kernel of the UNet plan is launched with cudaLaunchAttributeProgrammaticStreamSerialization and executes
// `griddepcontrol.launch_dependents` + `griddepcontrol.wait` after its own set-up (barrier init, TMEM allocation,
// constant-weight staging) and before its first access to global memory produced by a predecessor, so the ~330
// launches of one UNet call overlap their prologues with the predecessor's tail (also inside the captured graph).
bool use_tc_attention();
bool cluster_allowed();
void set_cluster_allowed(bool on);
bool pdl_enabled();
void set_pdl_enabled(bool on);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  int na = 1;
  if (cluster_x > 1) {
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = cluster_x;
    at[1].val.clusterDim.y = 1;
    at[1].val.clusterDim.z = 1;
    na = 2;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  return launch_kc(kernel, grid, block, smem, s, 1, static_cast<Args&&>(args)...);
}
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

// ------------------------------------------------------------------ GEMM / implicit-GEMM convolution
// D[M,N] = sum_seg A_seg[M,K_seg] . Wt[N,Ktot]^T  (+bias +temb +residual | GEGLU), fp16 operands, fp32 accumulate
// in TMEM.  A segments: segment 0 is either a plain [M,K] matrix (linear) or an NHWC activation read through
// `taps0` (1 or 9) shifted TMA boxes with zero fill at the borders (3x3 pad-1 convolution as implicit GEMM);
// segments 1,2 are extra 1-tap sources appended along K (fused 1x1 shortcut over a skip-concat).
struct ASource {
  const __half* ptr = nullptr;  // [rows, ld] fp16, `C` channels used
  int C = 0;
  int ld = 0;
};

struct alignas(64) GemmParams {
  CUtensorMap map_a[3];
  CUtensorMap map_b;
  CUtensorMap map_b_half;  // box of BN/2 weight rows (pair mode: each CTA of the pair stages half of the weight tile)
  int taps0, chunks0, chunks1, chunks2;
  int num_kb;
  int linear;
  int W, HW;
  int M, N;
  int m_tiles, n_tiles;
  const float* bias;        // [N] (packed order) or null
  const float* temb_table;  // per-timestep additive vector table or null
  const int* t_index;       // device pointer to the current row of temb_table
  int temb_stride;          // floats between rows of temb_table
  const __half* residual;   // [M, ldr] or null
  int ldr;
  __half* out;  // [M, ldc]
  int ldc;
  int geglu;  // epilogue: out[:, j] = (v_j) * gelu(g_j), tile columns [0,BN/2) = v, [BN/2,BN) = g
  // split-K (small-M, weight-streaming layers): `splits` CTAs share one output tile, each reduces a K range into
  // ws[split][M][N] (fp32); the last CTA to arrive (counters[tile]) sums the slices in index order and runs the epilogue
  int splits, kb_per_split;
  int raster;  // tile order of the persistent loop: 0 = M tiles fastest (CTAs in flight share a weight tile), 1 = N tiles fastest
               // (they share an activation tile: for activations larger than L2, which would be streamed once per N tile)
  float* ws;
  int* counters;
  volatile unsigned int* dbg;
  // optional per-CTA cycle counters [grid][8]: MMA thread (total, wait full, wait tempty), producer (total, wait empty),
  // epilogue thread 128 (total, wait tfull); null in production
  long long* prof;
  float* out32;  // GemmEpilogue::out_f32_nchw4
  int out32_ch;  // planes written there (first out32_ch columns of the tile)
  float out_scale;  // accumulators are multiplied by this before bias / residual (1 = off)
  const int* a0_row_map;  // conv mode: batch row the taps of A source 0 are read from, per output batch row (null = same)
  int exp;  // experiments (PNP_GEMM_EXP, test entry points only): 1 = no TMA copies, 2 = no MMAs
};

struct GemmPlan {
  GemmParams p;
  int bn = 0;    // N of one MMA
  int nsub = 1;  // accumulators per CTA tile (tile = 128 x nsub*bn)
  int cluster = 1;
  int grid = 0;
  size_t smem = 0;
  size_t algo_bytes = 0;  // compulsory HBM traffic of one launch: activations + weights + residual read once, output written once
};

struct GemmEpilogue {
  const float* bias = nullptr;
  const float* temb_table = nullptr;
  const int* t_index = nullptr;
  int temb_stride = 0;
  const __half* residual = nullptr;
  int ldr = 0;
  __half* out = nullptr;
  int ldc = 0;
  bool geglu = false;
  // conv_out: instead of the fp16 NHWC tile, columns 0..3 are written as fp32 NCHW [B,4,H,W] (the UNet's eps output);
  // the weight matrix is zero-padded to one 64-column tile
  float* out_f32_nchw4 = nullptr;
  int out32_channels = 4;  // ... or the first 1..8 columns (VAE: 3 image planes, 8 posterior moments)
  float out_scale = 1.0f;  // D = out_scale * (A . W^T) + bias ... (single-head VAE attention: 1/sqrt(C) on Q.K^T)
  // Plug-and-Play feature injection: source 0 (the 3x3 taps) of output batch row b is read from batch row a0_row_map[b]
  // (device array), the extra 1x1 shortcut sources stay row b's own.  Needs tiles that do not span images.
  const int* a0_row_map = nullptr;
};

// fp16 tiled tensor map with 128-byte swizzle and zero out-of-bounds fill (rank 2..4); strides in bytes for dims 1..
int encode_tensor_map_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box);

// conv-mode: srcs[0] is NHWC [B,H,W,C0] with `taps0`=9 (3x3, pad 1) or 1; linear mode: B=H=1, W=M.
int gemm_plan_create(GemmPlan* plan, const ASource* srcs, int nsrc, int taps0, bool linear, int B, int H, int W,
                     const __half* Wt, int N, int Ktot, const GemmEpilogue& ep, int bn_force, int num_sms,
                     int split_force = 0);
int gemm_launch(const GemmPlan& plan, cudaStream_t stream);
int gemm_choose_bn(int M, int N, bool geglu, int num_sms);
void gemm_choose(int M, int N, int num_kb, bool geglu, int num_sms, int bn_force, int split_force, int* bn_out,
                 int* splits_out);
// shared split-K workspace (all GEMMs of a stream run back to back): floats / ints needed by a plan
size_t gemm_ws_floats(const GemmPlan& plan);
void gemm_set_workspace(GemmPlan* plan, float* ws, int* counters);
constexpr int kGemmMaxCounters = 4096;
// modelled cycles of one launch (tile = bnt columns: 64/128/160/256, or 320 = two accumulators of 160); < 0 = invalid
long gemm_model_cost(int M, int N, int num_kb, bool geglu, int bnt, int splits, int num_sms);

// ------------------------------------------------------------------ normalisation kernels (norm.cu)
// GroupNorm(32 groups) [+SiLU] over NHWC fp16; optional second source = channel concat (skip connection).
int groupnorm_launch(const __half* x0, int C0, const __half* x1, int C1, int B, int HW, const float* gamma,
                     const float* beta, float eps, bool silu, __half* out, float* partials, cudaStream_t s);
size_t groupnorm_partials_floats(int B, int HW);
int groupnorm_path(int C, int B, int HW);  // 0 = statistics + apply, 1 = cluster kernel, 2 = register-resident kernel
int groupnorm_kernel_count(int C, int B, int HW);  // kernels groupnorm_launch will launch: 1 = register-resident or cluster kernel, 2 = statistics + apply
size_t groupnorm_workspace_floats(int B, int HW);  // partials + mean/rstd + per-batch counters (zero-initialised)
int layernorm_launch(const __half* x, int rows, int C, const float* gamma, const float* beta, float eps, __half* out,
                     cudaStream_t s);
int upsample2x_launch(const __half* x, int B, int H, int W, int C, __half* out, cudaStream_t s);
int im2col_s2_launch(const __half* x, int B, int H, int W, int C, __half* out, cudaStream_t s);
int conv_in_launch(const float* x_nchw, int B, int H, int W, const float* w, const float* bias, __half* out,
                   cudaStream_t s);
int concat_launch(const __half* x0, int C0, const __half* x1, int C1, int rows, __half* out, cudaStream_t s);

}  // namespace pnp
