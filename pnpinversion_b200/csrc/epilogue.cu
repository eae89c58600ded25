// Fused per-step latent update ("K10") and LocalBlend.  Pure HBM/launch-bound work on 64 KB per latent: one launch
// replaces the ~10 eager kernels of CFG + DDIM step + offset/rectification (+ ~10 more for LocalBlend) the reference
// issues per step (SURVEY.md section 8, rows a1,a2,a4,a5,a10).
#include "pnp_attn.h"
#include "pnp_internal.h"

namespace pnp {
namespace {

constexpr int LAT = 4 * 64 * 64;

__global__ void step_epilogue_kernel(const StepParams p) {
  const int total = p.n * (LAT / 4);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / (LAT / 4), e = i - r * (LAT / 4);
    const float4 x = reinterpret_cast<const float4*>(p.x)[i];
    float4 ec = reinterpret_cast<const float4*>(p.eps_c)[i];
    float eps[4] = {ec.x, ec.y, ec.z, ec.w};
    if (p.eps_u != nullptr) {
      const float4 eu = reinterpret_cast<const float4*>(p.eps_u)[i];
      const float u[4] = {eu.x, eu.y, eu.z, eu.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) eps[k] = __fadd_rn(u[k], __fmul_rn(p.guidance, __fsub_rn(eps[k], u[k])));
    }
    const float xs[4] = {x.x, x.y, x.z, x.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // explicit _rn intrinsics: no FMA contraction, same rounding sequence as the eager fp32 ops of the reference
      const float x0 = __fdiv_rn(__fsub_rn(xs[k], __fmul_rn(p.sqrt_1m_a_from, eps[k])), p.sqrt_a_from);
      o[k] = __fadd_rn(__fmul_rn(p.sqrt_a_to, x0), __fmul_rn(p.sqrt_1m_a_to, eps[k]));
    }
    if (p.target != nullptr) {
      const float4 tg = reinterpret_cast<const float4*>(p.target)[(r % p.target_rows) * (LAT / 4) + e];
      const float t[4] = {tg.x, tg.y, tg.z, tg.w};
      float l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        l[k] = __fsub_rn(t[k], o[k]);
        // ablations: `loss * scale` (inversion.py:489), zeros on skipped steps (:514-517); scale 1 multiplies exactly
        if (p.loss_scale != 1.0f) l[k] = __fmul_rn(l[k], p.loss_scale);
        o[k] = __fadd_rn(o[k], l[k]);
      }
      reinterpret_cast<float4*>(p.loss_out)[i] = make_float4(l[0], l[1], l[2], l[3]);
    }
    if (p.noise_loss != nullptr && ((p.add_mask >> r) & 1u)) {
      const float4 nl = reinterpret_cast<const float4*>(p.noise_loss)[i];
      o[0] = __fadd_rn(o[0], nl.x);
      o[1] = __fadd_rn(o[1], nl.y);
      o[2] = __fadd_rn(o[2], nl.z);
      o[3] = __fadd_rn(o[3], nl.w);
    }
    reinterpret_cast<float4*>(p.x_out)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// attention_control.py:97-121.  One CTA per (source, target) latent pair, thread = one 16x16 position.
//   get_mask(maps, alpha, use_pool): (maps * alpha).sum(-1).mean(1) [-> 3x3 max pool] -> nearest x4 -> / max -> > th
//   mask = get_mask(words, pool) ; if substruct: mask &= ~get_mask(sub_words, no pool) ; mask = mask[:1] + mask
__global__ void __launch_bounds__(256) local_blend_kernel(const LocalBlendParams p) {
  __shared__ float m[2][256];
  __shared__ float mp[2][256];
  __shared__ float red[2][8];
  __shared__ unsigned char mk[2][256];
  __shared__ unsigned char sk[2][256];
  const LocalBlendItem& it = p.items[blockIdx.x];
  const int pos = threadIdx.x;
  const int slot_of[2] = {it.src_slot, it.tgt_slot};
  const int y = pos / 16, x = pos % 16;
  for (int pass = 0; pass < 2; ++pass) {  // 0: blend words (pooled), 1: substruct words (not pooled)
    const bool sub = pass == 1;
    if (sub && it.nsub[0] == 0 && it.nsub[1] == 0) {
      sk[0][pos] = sk[1][pos] = 0;
      break;
    }
    for (int pr = 0; pr < 2; ++pr) {
      const int nw = sub ? it.nsub[pr] : it.nwords[pr];
      float acc = 0.f;
      for (int lh = 0; lh < 40; ++lh) {
        const int layer = lh / 8, h = lh % 8;
        const float* row =
            p.store + layer * p.layer_stride + slot_of[pr] * p.slot_stride + (static_cast<size_t>(h) * 256 + pos) * 77;
        float s = 0.f;
        for (int w = 0; w < nw; ++w)
          s += sub ? row[it.sub_words[pr][w]] * it.sub_alpha[pr][w] : row[it.words[pr][w]] * it.alpha[pr][w];
        acc += s;
      }
      m[pr][pos] = acc / 40.f;
    }
    __syncthreads();
    for (int pr = 0; pr < 2; ++pr) {
      float v;
      if (!sub) {
        v = -INFINITY;
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy >= 0 && yy < 16 && xx >= 0 && xx < 16) v = fmaxf(v, m[pr][yy * 16 + xx]);
          }
      } else {
        v = m[pr][pos];
      }
      mp[pr][pos] = v;
      float w = v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, o));
      if ((pos & 31) == 0) red[pr][pos >> 5] = w;
    }
    __syncthreads();
    for (int pr = 0; pr < 2; ++pr) {
      float mx = red[pr][0];
      for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[pr][i]);
      const unsigned char bit = (mp[pr][pos] / mx) > (sub ? it.th_sub : it.th_pool) ? 1 : 0;
      if (sub) sk[pr][pos] = bit; else mk[pr][pos] = bit;
    }
    __syncthreads();
  }
  __syncthreads();
  // mask = mask[:1] + mask (both the blend and the substruct mask) ; x_t = x_t[:1] + mask * (x_t - x_t[:1])
  float* xs = p.x + static_cast<size_t>(it.src_row) * LAT;
  float* xt = p.x + static_cast<size_t>(it.tgt_row) * LAT;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int yy = i / 64, xx = i % 64;
    const int cell = (yy / 4) * 16 + (xx / 4);
    const float m0 = (mk[0][cell] && !sk[0][cell]) ? 1.f : 0.f;
    const float m1 = ((mk[0][cell] | mk[1][cell]) && !(sk[0][cell] | sk[1][cell])) ? 1.f : 0.f;
    if (p.mask_out != nullptr) {
      p.mask_out[(static_cast<size_t>(blockIdx.x) * 2) * 4096 + i] = m0;
      p.mask_out[(static_cast<size_t>(blockIdx.x) * 2 + 1) * 4096 + i] = m1;
    }
    for (int c = 0; c < 4; ++c) {
      const float x0 = xs[c * 4096 + i];
      const float x1 = xt[c * 4096 + i];
      xs[c * 4096 + i] = __fadd_rn(x0, __fmul_rn(m0, __fsub_rn(x0, x0)));
      xt[c * 4096 + i] = __fadd_rn(x0, __fmul_rn(m1, __fsub_rn(x1, x0)));
    }
  }
}

// EDICT mixing layers (models/edict/edict_functions.py:854-859 reverse, :931-936 forward), in place on the coupled pair
__global__ void edict_mix_kernel(float* __restrict__ x, float* __restrict__ y, int n4, float w, int reverse) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(x)[i];
    float4 b = reinterpret_cast<float4*>(y)[i];
    float xs[4] = {a.x, a.y, a.z, a.w}, ys[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (reverse) {
        ys[k] = __fdiv_rn(__fsub_rn(ys[k], __fmul_rn(1.f - w, xs[k])), w);
        xs[k] = __fdiv_rn(__fsub_rn(xs[k], __fmul_rn(1.f - w, ys[k])), w);
      } else {
        xs[k] = __fadd_rn(__fmul_rn(w, xs[k]), __fmul_rn(1.f - w, ys[k]));
        ys[k] = __fadd_rn(__fmul_rn(1.f - w, xs[k]), __fmul_rn(w, ys[k]));
      }
    }
    reinterpret_cast<float4*>(x)[i] = make_float4(xs[0], xs[1], xs[2], xs[3]);
    reinterpret_cast<float4*>(y)[i] = make_float4(ys[0], ys[1], ys[2], ys[3]);
  }
}

}  // namespace

int edict_mix_launch(float* x, float* y, int n_elems, float w, bool reverse, cudaStream_t s) {
  PNP_CHECK(n_elems % 4 == 0 && x != nullptr && y != nullptr, "edict mix: bad arguments");
  edict_mix_kernel<<<(n_elems / 4 + 255) / 256, 256, 0, s>>>(x, y, n_elems / 4, w, reverse ? 1 : 0);
  PNP_CUDA(cudaGetLastError());
  return 0;
}

int step_epilogue_launch(const StepParams& p, cudaStream_t s) {
  PNP_CHECK(p.n >= 1 && p.n <= 32, "step epilogue: 1..32 latent rows");
  PNP_CHECK(p.target == nullptr || (p.loss_out != nullptr && p.target_rows >= 1), "step epilogue: offset mode");
  const int total = p.n * (LAT / 4);
  step_epilogue_kernel<<<(total + 255) / 256, 256, 0, s>>>(p);
  PNP_CUDA(cudaGetLastError());
  return 0;
}

int local_blend_launch(const LocalBlendParams& p, cudaStream_t s) {
  PNP_CHECK(p.n_items >= 1 && p.n_items <= 8, "local blend: 1..8 latent pairs per launch");
  for (int i = 0; i < p.n_items; ++i)
    for (int pr = 0; pr < 2; ++pr)
      PNP_CHECK(p.items[i].nwords[pr] >= 0 && p.items[i].nwords[pr] <= 8 && p.items[i].nsub[pr] >= 0 &&
                    p.items[i].nsub[pr] <= 8,
                "local blend: <= 8 words");
  local_blend_kernel<<<p.n_items, 256, 0, s>>>(p);
  PNP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pnp
