// HBM-bound helper kernels around the tcgen05 GEMM: GroupNorm(+SiLU) over NHWC, LayerNorm, nearest 2x upsample,
// stride-2 im2col, and the two tiny edge convolutions (4->320 and 320->4) that are not worth a tensor-core tile.
// All loads/stores are 16-byte vectors along the contiguous channel dimension.
//
// Reference arithmetic: torch.nn.GroupNorm(32, C, eps) / SiLU in models/edict/my_diffusers/models/resnet.py:336-358,
// GroupNorm(eps=1e-6) in attention.py:123,143; LayerNorm attention.py:195-200; Upsample2D resnet.py:38-52;
// Downsample2D (stride 2, pad 1) resnet.py:88-97; conv_in / conv_norm_out+conv_out unet_2d_condition.py:230,266-268.
#include <algorithm>

#include "pnp_internal.h"
#include "pnp_ptx.cuh"

namespace pnp {
namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __half22float2(h[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
  return u;
}
// x * sigmoid(x) with ex2.approx + rcp.approx (2 MUFU, ~2 ulp of fp32: far below the fp16 rounding of the result); an IEEE
// division here made the streaming normalisation passes issue-bound (ncu: 2.0 TB/s at 55 % issue utilisation).
// x -> -inf: exp = inf, 1 + inf = inf, __fdividef(x, inf) = -0.
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// ------------------------------------------------------------------ GroupNorm
constexpr int GN_GROUPS = 32;
constexpr int GN_MAX_VPT = 2;

// partials[b][slice][g] = (mean, M2) over ppc*cpg elements.  VPT = 16-byte vectors per thread (2 only for C > 2048): as a
// run-time value it cost 90 registers = 2 CTAs per SM = 30 KB of loads in flight per SM (ncu: 42 us for 84 MB).
template <int VPT>
__global__ void gn_stats_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                                int ppc, int tx_n, int rows_y, int vpt, float* __restrict__ partials, float eps,
                                float* __restrict__ mean_rstd, int* __restrict__ counters) {
  extern __shared__ float sm[];  // [rows_y][2*C] then reused
  pdl_sync();
  const int C = C0 + C1;
  const int nvec0 = C0 >> 3;
  const int b = blockIdx.y, slice = blockIdx.x, nslices = gridDim.x;
  const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;  // blockDim is always 256; rows ty >= rows_y idle
  float s[VPT][8], ss[VPT][8];
#pragma unroll
  for (int i = 0; i < VPT; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) s[i][e] = ss[i][e] = 0.f;
  if (ty < rows_y) {
    constexpr int U = 4;  // pixels in flight per thread (memory-level parallelism: these kernels are latency-bound)
    for (int pix0 = ty; pix0 < ppc; pix0 += U * rows_y) {
      uint4 u[U][VPT];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int pix = pix0 + j * rows_y;
        const size_t row = static_cast<size_t>(b) * HW + static_cast<size_t>(slice) * ppc + pix;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          u[j][i] = make_uint4(0, 0, 0, 0);
          if (i < vpt && pix < ppc) {
            const int v = tx + i * tx_n;
            u[j][i] = (v < nvec0) ? __ldg(reinterpret_cast<const uint4*>(x0 + row * C0 + v * 8))
                                  : __ldg(reinterpret_cast<const uint4*>(x1 + row * C1 + (v - nvec0) * 8));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          if (i < vpt) {
            float f[8];
            unpack8(u[j][i], f);  // out-of-range pixels were loaded as zeros: they add nothing
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              s[i][e] += f[e];
              ss[i][e] += f[e] * f[e];
            }
          }
        }
      }
    }
    float* mine = sm + static_cast<size_t>(ty) * 2 * C;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      if (i < vpt) {
        const int v = tx + i * tx_n;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          mine[v * 8 + e] = s[i][e];
          mine[C + v * 8 + e] = ss[i][e];
        }
      }
    }
  }
  __syncthreads();
  // reduce over ty into row 0
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float a = sm[c];
    for (int r = 1; r < rows_y; ++r) a += sm[static_cast<size_t>(r) * 2 * C + c];
    sm[c] = a;
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {
    const int g = threadIdx.x, cpg = C / GN_GROUPS;
    float a = 0.f, q = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      a += sm[c];
      q += sm[C + c];
    }
    const float n = static_cast<float>(ppc) * cpg;
    const float mean = a / n;
    const float m2 = fmaxf(q - a * mean, 0.f);
    float* o = partials + ((static_cast<size_t>(b) * nslices + slice) * GN_GROUPS + g) * 2;
    o[0] = mean;
    o[1] = m2;
  }
  // last CTA of this batch row merges the per-slice (mean, M2) pairs (Chan et al., equal counts) into (mean, rstd)
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(&counters[b], 1);
    is_last = (done == nslices - 1);
    if (is_last) counters[b] = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  {
    // 8 lanes per group
    const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
    if (g < GN_GROUPS) {
      const int cpg = C / GN_GROUPS;
      const float* pp = partials + (static_cast<size_t>(b) * nslices * GN_GROUPS + g) * 2;
      // all (mean, M2) pairs of this lane are fetched up front (one memory latency instead of a dependent chain)
      constexpr int MAXS = 16;  // nslices <= 128
      float2 pr[MAXS];
#pragma unroll
      for (int j = 0; j < MAXS; ++j) {
        const int s2 = sub + j * 8;
        pr[j] = make_float2(0.f, 0.f);
        if (s2 < nslices) pr[j] = __ldcg(reinterpret_cast<const float2*>(pp + static_cast<size_t>(s2) * GN_GROUPS * 2));
      }
      float msum = 0.f;
#pragma unroll
      for (int j = 0; j < MAXS; ++j) msum += pr[j].x;  // missing slices contribute 0
      msum += __shfl_xor_sync(0xffffffffu, msum, 1);
      msum += __shfl_xor_sync(0xffffffffu, msum, 2);
      msum += __shfl_xor_sync(0xffffffffu, msum, 4);
      const float mean = msum / nslices;
      const float n_i = static_cast<float>(ppc) * cpg;
      float m2 = 0.f;
#pragma unroll
      for (int j = 0; j < MAXS; ++j) {
        if (sub + j * 8 < nslices) {
          const float d = pr[j].x - mean;
          m2 += pr[j].y + n_i * d * d;
        }
      }
      m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
      m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
      m2 += __shfl_xor_sync(0xffffffffu, m2, 4);
      if (sub == 0) {
        mean_rstd[(b * GN_GROUPS + g) * 2] = mean;
        mean_rstd[(b * GN_GROUPS + g) * 2 + 1] = rsqrtf(m2 / (n_i * nslices) + eps);
      }
    }
  }
}

__global__ void gn_apply_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                                const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int do_silu, __half* __restrict__ out, int ppc) {
  extern __shared__ float sm[];  // scale[C], shift[C], mean[32], rstd[32]
  pdl_sync();
  const int C = C0 + C1;
  float* scale = sm;
  float* shift = sm + C;
  float* gmean = sm + 2 * C;
  float* grstd = gmean + GN_GROUPS;
  const int b = blockIdx.y;
  const int cpg = C / GN_GROUPS;
  if (threadIdx.x < GN_GROUPS) {
    gmean[threadIdx.x] = mean_rstd[(b * GN_GROUPS + threadIdx.x) * 2];
    grstd[threadIdx.x] = mean_rstd[(b * GN_GROUPS + threadIdx.x) * 2 + 1];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float sc = grstd[g] * gamma[c];
    scale[c] = sc;
    shift[c] = beta[c] - gmean[g] * sc;
  }
  __syncthreads();
  const int nvec = C >> 3, nvec0 = C0 >> 3;
  const int total = ppc * nvec;
  const size_t row0 = static_cast<size_t>(b) * HW + static_cast<size_t>(blockIdx.x) * ppc;
  constexpr int U = 4;
  for (int base = threadIdx.x; base < total; base += U * blockDim.x) {
    uint4 u[U];
    int pixs[U], vs[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int idx = base + j * blockDim.x;
      const int pix = idx / nvec, v = idx - pix * nvec;
      pixs[j] = pix;
      vs[j] = v;
      if (idx < total) {
        const size_t row = row0 + pix;
        u[j] = (v < nvec0) ? __ldg(reinterpret_cast<const uint4*>(x0 + row * C0 + v * 8))
                           : __ldg(reinterpret_cast<const uint4*>(x1 + row * C1 + (v - nvec0) * 8));
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int idx = base + j * blockDim.x;
      if (idx < total) {
        float f[8];
        unpack8(u[j], f);
        const int v = vs[j];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float y = f[e] * scale[v * 8 + e] + shift[v * 8 + e];
          f[e] = do_silu ? silu(y) : y;
        }
        *reinterpret_cast<uint4*>(out + (row0 + pixs[j]) * C + v * 8) = pack8(f);
      }
    }
  }
}

// Streaming normalise (+SiLU), the HBM-bound half of GroupNorm for image batches (84 MB in + 84 MB out per launch at
// B = 32, 64x64x320).  The block size is a multiple of the 16-byte vectors per pixel, so the flat vector index
// tid + k * blockDim keeps its channel vector: every thread owns ONE channel vector, its eight scale / shift pairs live in
// registers, and the loop is U independent 16-byte loads, 8 FMAs + SiLUs each, U stores - no integer division, no shared
// memory.  `reverse`: CTAs walk the tensor back to front.  The statistics pass (and the GEMM before it) touched the
// tensor front to back, so with an LRU-like L2 smaller than the tensor a front-to-back reader misses everything while a
// back-to-front reader still finds the tail; and the consumer (an implicit-GEMM conv reading front to back) then finds
// the head of the output this kernel wrote last.
template <int U>
__global__ void __launch_bounds__(512) gn_apply_vec_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1,
                                                           int C1, int HW, const float* __restrict__ mean_rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int do_silu, __half* __restrict__ out, int ppc, int reverse) {
  pdl_sync();
  const int C = C0 + C1, nvec = C >> 3, nvec0 = C0 >> 3;
  const int b = reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int slice = reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int v = threadIdx.x % nvec, py = threadIdx.x / nvec, rows_y = blockDim.x / nvec;
  const int cpg = C / GN_GROUPS;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = v * 8 + e, g = c / cpg;
    const float2 mr = __ldg(reinterpret_cast<const float2*>(mean_rstd + (b * GN_GROUPS + g) * 2));
    sc[e] = mr.y * __ldg(gamma + c);
    sh[e] = __ldg(beta + c) - mr.x * sc[e];
  }
  const size_t row0 = static_cast<size_t>(b) * HW + static_cast<size_t>(slice) * ppc;
  const bool first = v < nvec0;
  const __half* src = first ? x0 + row0 * C0 + v * 8 : x1 + row0 * C1 + (v - nvec0) * 8;
  const size_t sstride = first ? C0 : C1;
  __half* dst = out + row0 * C + v * 8;
  for (int p0 = py; p0 < ppc; p0 += U * rows_y) {
    uint4 u[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int p = p0 + j * rows_y;
      if (p < ppc) u[j] = __ldg(reinterpret_cast<const uint4*>(src + p * sstride));
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int p = p0 + j * rows_y;
      if (p < ppc) {
        float f[8];
        unpack8(u[j], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float y = fmaf(f[e], sc[e], sh[e]);
          f[e] = do_silu ? silu(y) : y;
        }
        *reinterpret_cast<uint4*>(dst + static_cast<size_t>(p) * C) = pack8(f);
      }
    }
  }
}

// Register-resident GroupNorm for the low-resolution levels of an image batch.  One CTA owns one image x one chunk of G
// whole groups (Cw = G * C/32 channels = W16 16-byte vectors per pixel): its HW x Cw halves (<= 88 KB) are loaded ONCE,
// all loads in flight at once, and stay in registers while the CTA computes the mean, then the centred second moment,
// then normalises (+SiLU) and stores.  No inter-CTA traffic, no second read, one launch.  With B x 32/G >= 148 CTAs this
// replaces the 16-CTA cluster kernel (two clusters per GPC at a time: 43-58 us per launch at B = 32 for L2-resident
// tensors) and the statistics + apply pair of the 32x32 level (22-29 + 21-25 us).
template <int NV>  // 16-byte vectors per thread; the 12-vector variant runs 448 threads (146 registers each, no spills)
__global__ void __launch_bounds__(NV > 8 ? 448 : 512) gn_local_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1,
                                                       int C1, int HW, int W16, int rows_y, float eps,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int do_silu, __half* __restrict__ out) {
  extern __shared__ float sm[];  // red[rows_y][Cw] | part[<= blockDim] | col[Cw] | gstat[G]
  pdl_sync();
  const int C = C0 + C1, nvec0 = C0 >> 3, cpg = C / GN_GROUPS;
  const int Cw = W16 * 8, G = Cw / cpg;
  float* red = sm;
  float* part = sm + static_cast<size_t>(rows_y) * Cw;
  float* col = part + blockDim.x;
  float* gstat = col + Cw;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int v = threadIdx.x % W16, py = threadIdx.x / W16;  // blockDim = rows_y * W16
  const int vg = chunk * W16 + v;                           // channel vector within the pixel
  const bool first = vg < nvec0;
  const size_t row0 = static_cast<size_t>(b) * HW;
  const __half* src = first ? x0 + row0 * C0 + vg * 8 : x1 + row0 * C1 + (vg - nvec0) * 8;
  const size_t sstride = first ? C0 : C1;
  uint4 u[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int p = py + k * rows_y;
    u[k] = make_uint4(0, 0, 0, 0);
    if (p < HW) u[k] = __ldg(reinterpret_cast<const uint4*>(src + p * sstride));
  }
  const float inv_n = 1.0f / (static_cast<float>(HW) * cpg);
  // CTA-wide reduction of eight per-thread channel sums to per-group values in gstat (rows -> channels -> groups)
  auto reduce_groups = [&](const float (&acc)[8], bool second) {
    float* mine = red + static_cast<size_t>(py) * Cw + v * 8;
    *reinterpret_cast<float4*>(mine) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(mine + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    __syncthreads();
    {
      // rows -> channels in two levels so that every thread takes part (a single level is a chain of rows_y loads)
      const int nseg = blockDim.x / Cw;  // >= 1: blockDim = rows_y * W16 >= Cw / 8 ... see the launcher (rows_y >= 8)
      const int c = threadIdx.x % Cw, seg = threadIdx.x / Cw;
      if (seg < nseg) {
        float a = 0.f;
        for (int r = seg; r < rows_y; r += nseg) a += red[static_cast<size_t>(r) * Cw + c];
        part[seg * Cw + c] = a;
      }
      __syncthreads();
      if (threadIdx.x < Cw) {
        float a = 0.f;
        for (int sgi = 0; sgi < nseg; ++sgi) a += part[sgi * Cw + threadIdx.x];
        col[threadIdx.x] = a;
      }
    }
    __syncthreads();
    if (threadIdx.x < G) {
      float a = 0.f;
      for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) a += col[c];
      gstat[threadIdx.x] = second ? rsqrtf(a * inv_n + eps) : a * inv_n;
    }
    __syncthreads();
  };
  float mean[8];
  {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {  // out-of-range pixels were loaded as zeros
      float f[8];
      unpack8(u[k], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
    reduce_groups(acc, false);
#pragma unroll
    for (int e = 0; e < 8; ++e) mean[e] = gstat[(v * 8 + e) / cpg];
    __syncthreads();  // gstat is rewritten by the second round
  }
  {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (py + k * rows_y < HW) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = f[e] - mean[e];
          acc[e] = fmaf(d, d, acc[e]);
        }
      }
    }
    reduce_groups(acc, true);
  }
  float sc[8], sh[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vg * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + vg * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vg * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + vg * 8 + 4));
    const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = gstat[(v * 8 + e) / cpg] * ga[e];
      sh[e] = be[e] - mean[e] * sc[e];
    }
  }
  __half* dst = out + row0 * C + vg * 8;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int p = py + k * rows_y;
    if (p < HW) {
      float f[8];
      unpack8(u[k], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float y = fmaf(f[e], sc[e], sh[e]);
        f[e] = do_silu ? silu(y) : y;
      }
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(p) * C) = pack8(f);
    }
  }
}

// geometry of the register-resident path for this shape (false = does not apply)
struct GnLocal {
  int G, W16, rows_y, nv;
};
static bool gn_local_for(int C, int B, int HW, GnLocal* o) {
  static const int enabled = [] { const char* e = getenv("PNP_GN_LOCAL"); return e ? atoi(e) : 1; }();
  if (!enabled) return false;
  const int cpg = C / GN_GROUPS;
  int best = 0;
  for (int G = 1; G <= 8; G <<= 1) {
    const int Cw = G * cpg;
    if (Cw % 8 != 0 || GN_GROUPS % G != 0) continue;
    const size_t bytes = static_cast<size_t>(HW) * Cw * 2;
    if (bytes > 88 * 1024) break;
    if (static_cast<long>(B) * (GN_GROUPS / G) < 148) break;  // too few CTAs: the cluster / two-kernel paths spread better
    best = G;  // the largest chunk that still fills the chip (longer contiguous runs per pixel)
    if (static_cast<long>(B) * (GN_GROUPS / G) < 2 * 256) break;
  }
  if (!best) return false;
  const int W16 = best * cpg / 8;
  int rows_y = std::min(HW, 512 / W16);
  int nv = (HW + rows_y - 1) / rows_y;
  if (nv > 8) {
    rows_y = std::min(HW, 448 / W16);
    nv = (HW + rows_y - 1) / rows_y;
  }
  if (rows_y < 8 || nv > 12) return false;  // rows_y >= 8: the block has at least Cw threads (two-level reduction)
  *o = GnLocal{best, W16, rows_y, nv};
  return true;
}

// Single-launch GroupNorm: one thread-block cluster per image.  Each CTA owns HW/CS pixels, reduces them to per-group
// (mean, M2), the CTAs exchange those 64 floats through distributed shared memory, and every CTA then normalises its
// own pixels (second read is an L2 hit).  Replaces stats + atomics/last-block + apply: the dependent chain is
// load -> block reduce -> cluster barrier -> load -> store, with no global round trip for the statistics.
// 512 threads with 8 (VPT=1) or 4 (VPT=2) 16-byte loads in flight each: one SM sustains bytes-in-flight / L2 latency, and
// with only 16 CTAs per image the first version (256 threads x 4 loads) ran at 10 GB/s per SM.
constexpr int GNC_THREADS = 512;
template <int VPT, int U>
__global__ void __launch_bounds__(GNC_THREADS) gn_cluster_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                                  int tx_n, int rows_y, float eps, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, int do_silu, __half* __restrict__ out) {
  extern __shared__ float sm[];  // [rows_y][2*C] reduction scratch, then scale[C] | shift[C]
  __shared__ __align__(8) float part[GN_GROUPS * 2];
  __shared__ float gstat[GN_GROUPS * 2];
  pdl_sync();
  uint32_t CS;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(CS));
  const uint32_t rank = cluster_ctarank();
  const int C = C0 + C1;
  const int nvec = C >> 3, nvec0 = C0 >> 3;
  const int b = blockIdx.y;
  const int ppc = HW / static_cast<int>(CS);
  const size_t row0 = static_cast<size_t>(b) * HW + static_cast<size_t>(rank) * ppc;
  const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
  const int cpg = C / GN_GROUPS;
  float s[VPT][8], ss[VPT][8];
#pragma unroll
  for (int i = 0; i < VPT; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) s[i][e] = ss[i][e] = 0.f;
  if (ty < rows_y) {
    for (int pix0 = ty; pix0 < ppc; pix0 += U * rows_y) {
      uint4 u[U][VPT];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int pix = pix0 + j * rows_y;
        const size_t row = row0 + pix;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          u[j][i] = make_uint4(0, 0, 0, 0);
          if (pix < ppc) {
            const int v = tx + i * tx_n;
            u[j][i] = (v < nvec0) ? __ldg(reinterpret_cast<const uint4*>(x0 + row * C0 + v * 8))
                                  : __ldg(reinterpret_cast<const uint4*>(x1 + row * C1 + (v - nvec0) * 8));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          {
            float f[8];
            unpack8(u[j][i], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              s[i][e] += f[e];
              ss[i][e] += f[e] * f[e];
            }
          }
        }
      }
    }
    float* mine = sm + static_cast<size_t>(ty) * 2 * C;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      {
        const int v = tx + i * tx_n;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          mine[v * 8 + e] = s[i][e];
          mine[C + v * 8 + e] = ss[i][e];
        }
      }
    }
  }
  // affine parameters are fetched while the reduction runs (their latency must not sit behind the cluster barrier)
  constexpr int CPT = 5;  // C <= 2560 (checked by the launcher) -> at most 5 channels per thread
  float gam[CPT], bet[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int c = threadIdx.x + i * GNC_THREADS;
    gam[i] = c < C ? __ldg(gamma + c) : 0.f;
    bet[i] = c < C ? __ldg(beta + c) : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float a = sm[c];
    for (int r = 1; r < rows_y; ++r) a += sm[static_cast<size_t>(r) * 2 * C + c];
    sm[c] = a;
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {
    const int g = threadIdx.x;
    float a = 0.f, q = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      a += sm[c];
      q += sm[C + c];
    }
    const float n = static_cast<float>(ppc) * cpg;
    const float mean = a / n;
    part[g * 2] = mean;
    part[g * 2 + 1] = fmaxf(q - a * mean, 0.f);
  }
  cluster_sync_all();
  if (threadIdx.x < GN_GROUPS) {
    // Chan et al. merge with equal counts, same order in every CTA -> identical statistics cluster-wide
    const int g = threadIdx.x;
    const uint32_t local = smem_u32(&part[g * 2]);
    float2 pr[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      pr[r] = make_float2(0.f, 0.f);
      if (r < static_cast<int>(CS)) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
        asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(pr[r].x), "=f"(pr[r].y) : "r"(remote));
      }
    }
    float msum = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) msum += pr[r].x;
    const float mean = msum / static_cast<float>(CS);
    const float n_i = static_cast<float>(ppc) * cpg;
    float m2 = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (r < static_cast<int>(CS)) {
        const float d = pr[r].x - mean;
        m2 += pr[r].y + n_i * d * d;
      }
    }
    gstat[g * 2] = mean;
    gstat[g * 2 + 1] = rsqrtf(m2 / (n_i * static_cast<float>(CS)) + eps);
  }
  // peers may still be reading part[]: arrive now, wait just before exit
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  __syncthreads();
  float* scale = sm;
  float* shift = sm + C;
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int c = threadIdx.x + i * GNC_THREADS;
    if (c < C) {
      const int g = c / cpg;
      const float sc = gstat[g * 2 + 1] * gam[i];
      scale[c] = sc;
      shift[c] = bet[i] - gstat[g * 2] * sc;
    }
  }
  __syncthreads();
  {
    const int total = ppc * nvec;
    for (int base = threadIdx.x; base < total; base += U * blockDim.x) {
      uint4 u[U];
      int pixs[U], vs[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int idx = base + j * blockDim.x;
        const int pix = idx / nvec, v = idx - pix * nvec;
        pixs[j] = pix;
        vs[j] = v;
        if (idx < total) {
          const size_t row = row0 + pix;
          u[j] = (v < nvec0) ? __ldg(reinterpret_cast<const uint4*>(x0 + row * C0 + v * 8))
                             : __ldg(reinterpret_cast<const uint4*>(x1 + row * C1 + (v - nvec0) * 8));
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int idx = base + j * blockDim.x;
        if (idx < total) {
          float f[8];
          unpack8(u[j], f);
          const int v = vs[j];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float y = f[e] * scale[v * 8 + e] + shift[v * 8 + e];
            f[e] = do_silu ? silu(y) : y;
          }
          *reinterpret_cast<uint4*>(out + (row0 + pixs[j]) * C + v * 8) = pack8(f);
        }
      }
    }
  }
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int gn_ppc(int B, int HW) {
  // pixels per CTA: aim for >= ~256 CTAs (tuned at B <= 4), ~1024 for image batches (B >= 16: 84 MB tensors, the
  // statistics pass wants every SM several CTAs deep), at least 8 pixels each
  const long want = B >= 16 ? 1024 : 256;
  int ppc = HW;
  while (ppc > 8 && HW / ppc < 128 && static_cast<long>(B) * (HW / ppc) < want) ppc >>= 1;  // <= 128 slices
  return ppc;
}

// ------------------------------------------------------------------ LayerNorm: one warp per token
template <int VPL, int R>  // VPL 16-byte vectors per lane (C <= 8*32*VPL), R rows in flight per warp
__global__ void __launch_bounds__(256, VPL <= 2 ? 3 : 1)
ln_kernel(const __half* __restrict__ x, int rows, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
          float eps, __half* __restrict__ out) {
  // gamma / beta live in shared memory: as 16 * VPL registers per thread they held the kernel at 2 CTAs per SM (ncu: 3 TB/s)
  __shared__ __align__(16) float sgam[8 * 32 * VPL], sbet[8 * 32 * VPL];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    sgam[c] = gamma[c];
    sbet[c] = beta[c];
  }
  __syncthreads();
  pdl_sync();  // the parameters are constants: staged before waiting for the predecessor
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = C >> 3;
  for (int row0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * R; row0 < rows; row0 += nwarps * R) {
    uint4 u[R][VPL];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + i * 32;
        u[r][i] = make_uint4(0, 0, 0, 0);
        if (v < nvec && row0 + r < rows)
          u[r][i] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(row0 + r) * C + v * 8));
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r >= rows) break;
      float f[VPL][8];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        unpack8(u[r][i], f[i]);  // lanes beyond nvec hold zeros
#pragma unroll
        for (int e = 0; e < 8; ++e) s += f[i][e];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / C;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (lane + i * 32 < nvec) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float d = f[i][e] - mean;
            q += d * d;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q / C + eps);
      __half* orow = out + static_cast<size_t>(row0 + r) * C;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
          const float4 g0 = *reinterpret_cast<const float4*>(sgam + v * 8), g1 = *reinterpret_cast<const float4*>(sgam + v * 8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(sbet + v * 8), b1 = *reinterpret_cast<const float4*>(sbet + v * 8 + 4);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = (f[i][e] - mean) * rstd * gg[e] + bb[e];
          *reinterpret_cast<uint4*>(orow + v * 8) = pack8(y);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ data movement
__global__ void upsample2x_kernel(const uint4* __restrict__ x, int B, int H, int W, int nvec, uint4* __restrict__ out) {
  pdl_sync();
  const size_t total = static_cast<size_t>(B) * 2 * H * 2 * W * nvec;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = idx % nvec;
    size_t r = idx / nvec;
    const int xo = r % (2 * W);
    r /= (2 * W);
    const int yo = r % (2 * H);
    const int b = r / (2 * H);
    out[idx] = x[((static_cast<size_t>(b) * H + (yo >> 1)) * W + (xo >> 1)) * nvec + v];
  }
}

// out[(b,yo,xo)][tap*C + c] = x[b, 2yo+dy-1, 2xo+dx-1, c]  (zero outside), tap = dy*3+dx
__global__ void im2col_s2_kernel(const uint4* __restrict__ x, int B, int H, int W, int nvec, uint4* __restrict__ out) {
  pdl_sync();
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * 9 * nvec;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = idx % nvec;
    size_t r = idx / nvec;
    const int tap = r % 9;
    r /= 9;
    const int xo = r % Wo;
    r /= Wo;
    const int yo = r % Ho;
    const int b = r / Ho;
    const int yi = 2 * yo + tap / 3 - 1, xi = 2 * xo + tap % 3 - 1;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (yi >= 0 && yi < H && xi >= 0 && xi < W) u = x[((static_cast<size_t>(b) * H + yi) * W + xi) * nvec + v];
    out[idx] = u;
  }
}

// ------------------------------------------------------------------ conv_in: NCHW fp32 (4 ch) -> NHWC fp16 (320 ch)
constexpr int CIN_K = 36;
constexpr int CIN_CO = 320;
// One thread = 8 output channels x 8 consecutive pixels of an image row.  The [k][co] weights live in shared memory and
// every weight vector fetched serves 8 pixels (the first version fetched 2 x 16 bytes of weights per pixel and tap:
// 3 GB of shared-memory reads per B=4 call, 60 us); the 3 x 10 input window of a channel is read once per row of taps.
constexpr int CIN_PX = 8;
__global__ void __launch_bounds__(240) conv_in_kernel(const float* __restrict__ x, int B, int H, int W,
                                                      const float* __restrict__ w, const float* __restrict__ bias,
                                                      __half* __restrict__ out) {
  __shared__ __align__(16) float ws[CIN_K * CIN_CO];  // [k][co], k = ci*9 + tap
  for (int i = threadIdx.x; i < CIN_K * CIN_CO; i += blockDim.x) ws[i] = w[i];  // prepacked [k][co]
  __syncthreads();
  pdl_sync();  // the weights are constants: staged before waiting for the predecessor
  const int nvec = CIN_CO / 8;  // 40
  const int gpb = blockDim.x / nvec;  // pixel groups per CTA
  const int v = threadIdx.x % nvec, gl = threadIdx.x / nvec;
  const int gpr = W / CIN_PX;  // groups per image row
  const size_t ngroups = static_cast<size_t>(B) * H * gpr;
  float bs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) bs[e] = bias[v * 8 + e];
  for (size_t grp = static_cast<size_t>(blockIdx.x) * gpb + gl; grp < ngroups; grp += static_cast<size_t>(gridDim.x) * gpb) {
    const int x0 = static_cast<int>(grp % gpr) * CIN_PX;
    const int yy = static_cast<int>((grp / gpr) % H);
    const int b = static_cast<int>(grp / (static_cast<size_t>(gpr) * H));
    float acc[CIN_PX][8];
#pragma unroll
    for (int j = 0; j < CIN_PX; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[j][e] = bs[e];
    for (int ci = 0; ci < 4; ++ci) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yi = yy + dy - 1;
        float xin[CIN_PX + 2];
        const float* row = x + ((static_cast<size_t>(b) * 4 + ci) * H + yi) * W;
#pragma unroll
        for (int j = 0; j < CIN_PX + 2; ++j) {
          const int xi = x0 + j - 1;
          xin[j] = (yi >= 0 && yi < H && xi >= 0 && xi < W) ? __ldg(row + xi) : 0.f;
        }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float* wp = ws + (ci * 9 + dy * 3 + dx) * CIN_CO + v * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wp);
          const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
          for (int j = 0; j < CIN_PX; ++j) {
            const float a = xin[j + dx];
            acc[j][0] += a * w0.x; acc[j][1] += a * w0.y; acc[j][2] += a * w0.z; acc[j][3] += a * w0.w;
            acc[j][4] += a * w1.x; acc[j][5] += a * w1.y; acc[j][6] += a * w1.z; acc[j][7] += a * w1.w;
          }
        }
      }
    }
    const size_t pix0 = (static_cast<size_t>(b) * H + yy) * W + x0;
#pragma unroll
    for (int j = 0; j < CIN_PX; ++j) *reinterpret_cast<uint4*>(out + (pix0 + j) * CIN_CO + v * 8) = pack8(acc[j]);
  }
}

}  // namespace

size_t groupnorm_partials_floats(int B, int HW) {
  const int ppc = gn_ppc(B, HW);
  return static_cast<size_t>(B) * (HW / ppc) * GN_GROUPS * 2;
}

// Cluster size of the single-launch GroupNorm kernel this device can co-schedule (0 = unavailable / disabled) and whether it
// is forced for every shape.  Decided once per process.
static int gn_cluster_config(int* force_all) {
  static int cluster_cs = -1;
  static int cluster_force_all = 0;
  if (cluster_cs < 0) {
    cluster_cs = 0;
    const char* e = getenv("PNP_GN_CLUSTER");
    // measured on B200 (profiles/README.md): the cluster kernel wins on the small tensors (8x8: 10.5 vs 14.9 us, 16x16:
    // 11-15 vs 14-17 us) and loses on the large ones (64x64: 31 vs 23 us: 16 CTAs per image cannot keep enough loads in
    // flight), so by default it takes HW <= 256 only.  PNP_GN_CLUSTER=0 never, =8/16 always (tests).
    const int want = e ? atoi(e) : 16;
    cluster_force_all = e != nullptr && want >= 2;
    if (want >= 2) {
      bool ok = true;
      ok &= cudaFuncSetAttribute(gn_cluster_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) == cudaSuccess;
      ok &= cudaFuncSetAttribute(gn_cluster_kernel<1, 8>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
      ok &= cudaFuncSetAttribute(gn_cluster_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) == cudaSuccess;
      ok &= cudaFuncSetAttribute(gn_cluster_kernel<2, 4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
      if (!ok) (void)cudaGetLastError();  // no cluster kernel on this device: the two-kernel path serves every shape
      for (int cs = want; ok && cs >= 2 && !cluster_cs; cs >>= 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs, 1);
        cfg.blockDim = dim3(GNC_THREADS);
        cfg.dynamicSmemBytes = 96 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, gn_cluster_kernel<1, 8>, &cfg) == cudaSuccess && nclusters >= 4)
          cluster_cs = cs;
        else
          (void)cudaGetLastError();
      }
    }
  }
  *force_all = cluster_force_all;
  return cluster_cs;
}

// cluster size the launcher will use for this shape, 0 = two-kernel path
static int gn_cluster_for(int C, int HW) {
  int force_all = 0;
  const int cluster_cs = gn_cluster_config(&force_all);
  if (!(cluster_cs && C <= 2560 && (force_all || HW <= 256))) return 0;
  int cs = cluster_cs;
  while (cs > 1 && (HW % cs != 0 || HW / cs < 1)) cs >>= 1;
  return cs >= 2 ? cs : 0;
}

int groupnorm_path(int C, int B, int HW) {
  GnLocal g;
  if (gn_local_for(C, B, HW, &g)) return 2;
  return gn_cluster_for(C, HW) ? 1 : 0;
}
int groupnorm_kernel_count(int C, int B, int HW) { return groupnorm_path(C, B, HW) ? 1 : 2; }

int groupnorm_launch(const __half* x0, int C0, const __half* x1, int C1, int B, int HW, const float* gamma,
                     const float* beta, float eps, bool do_silu, __half* out, float* partials, cudaStream_t s) {
  const int C = C0 + C1;
  PNP_CHECK(C % (8 * GN_GROUPS / 8) == 0 && C % GN_GROUPS == 0 && C0 % 8 == 0 && C1 % 8 == 0, "groupnorm: channels");
  const int nvec = C / 8;
  const int vpt = (nvec + 255) / 256;
  PNP_CHECK(vpt <= GN_MAX_VPT && nvec % vpt == 0, "groupnorm: too many channels");
  const int tx_n = nvec / vpt;
  const int rows_y = 256 / tx_n;
  const int threads = tx_n * rows_y;
  const int ppc = gn_ppc(B, HW);
  const int nslices = HW / ppc;
  PNP_CHECK(HW % ppc == 0 && nslices <= 128, "groupnorm: HW split");
  const size_t sm1 = static_cast<size_t>(rows_y) * 2 * C * sizeof(float);
  static bool attr1 = false;
  if (!attr1) {
    PNP_CUDA(cudaFuncSetAttribute(gn_stats_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    PNP_CUDA(cudaFuncSetAttribute(gn_stats_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr1 = true;
  }
  PNP_CHECK(sm1 <= 160 * 1024, "groupnorm: smem");
  // workspace layout (fixed offsets so that the self-resetting counters are never overwritten by another shape):
  //   [0,64) ints: per-batch arrival counters | [64, 64+64*64) floats: (mean, rstd) | partials
  int* counters = reinterpret_cast<int*>(partials);
  float* mean_rstd = partials + 64;
  float* parts = partials + 64 + 64 * GN_GROUPS * 2;
  PNP_CHECK(B <= 64, "groupnorm: batch");
  (void)threads;
  GnLocal gl;
  if (gn_local_for(C, B, HW, &gl)) {
    const int Cw = gl.W16 * 8;
    const size_t sml = (static_cast<size_t>(gl.rows_y) * Cw + gl.rows_y * gl.W16 + Cw + 8) * sizeof(float);
    static bool attr_l = false;
    if (!attr_l) {
      PNP_CUDA(cudaFuncSetAttribute(gn_local_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      PNP_CUDA(cudaFuncSetAttribute(gn_local_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      PNP_CUDA(cudaFuncSetAttribute(gn_local_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      attr_l = true;
    }
    PNP_CHECK(sml <= 64 * 1024, "groupnorm: local scratch");
    const dim3 grid(GN_GROUPS / gl.G, B), block(gl.rows_y * gl.W16);
    if (gl.nv <= 4)
      PNP_CUDA(launch_k(gn_local_kernel<4>, grid, block, sml, s, x0, C0, x1, C1, HW, gl.W16, gl.rows_y, eps, gamma, beta,
                        do_silu ? 1 : 0, out));
    else if (gl.nv <= 8)
      PNP_CUDA(launch_k(gn_local_kernel<8>, grid, block, sml, s, x0, C0, x1, C1, HW, gl.W16, gl.rows_y, eps, gamma, beta,
                        do_silu ? 1 : 0, out));
    else
      PNP_CUDA(launch_k(gn_local_kernel<12>, grid, block, sml, s, x0, C0, x1, C1, HW, gl.W16, gl.rows_y, eps, gamma, beta,
                        do_silu ? 1 : 0, out));
    return 0;
  }
  // single-launch cluster kernel (one cluster of 16 or 8 CTAs per image) when the device can co-schedule it
  if (const int cs = gn_cluster_for(C, HW)) {
    const int ppcc = HW / cs;
    int ry = GNC_THREADS / tx_n;  // pixel rows of threads; no more than there are pixels, and the scratch must fit
    ry = std::max(1, std::min(ry, std::min(ppcc, static_cast<int>((96 * 1024) / (2 * C * sizeof(float))))));
    const size_t smc = static_cast<size_t>(ry) * 2 * C * sizeof(float);
    if (vpt == 1)
      PNP_CUDA(launch_kc(gn_cluster_kernel<1, 8>, dim3(cs, B), dim3(GNC_THREADS), smc, s, cs, x0, C0, x1, C1, HW, tx_n, ry,
                         eps, gamma, beta, do_silu ? 1 : 0, out));
    else
      PNP_CUDA(launch_kc(gn_cluster_kernel<2, 4>, dim3(cs, B), dim3(GNC_THREADS), smc, s, cs, x0, C0, x1, C1, HW, tx_n, ry,
                         eps, gamma, beta, do_silu ? 1 : 0, out));
    return 0;
  }
  if (vpt == 1)
    PNP_CUDA(launch_k(gn_stats_kernel<1>, dim3(nslices, B), dim3(256), sm1, s, x0, C0, x1, C1, HW, ppc, tx_n, rows_y, vpt,
                      parts, eps, mean_rstd, counters));
  else
    PNP_CUDA(launch_k(gn_stats_kernel<2>, dim3(nslices, B), dim3(256), sm1, s, x0, C0, x1, C1, HW, ppc, tx_n, rows_y, vpt,
                      parts, eps, mean_rstd, counters));
  const size_t sm2 = (2 * static_cast<size_t>(C) + 2 * GN_GROUPS) * sizeof(float);
  // the apply pass is pure streaming: fewer, fatter CTAs than the statistics pass
  // one pass of 4 vectors per thread per CTA where possible (1024 vectors per CTA)
  // ... up to ~8 CTAs per SM; beyond that (image batches: B >= 16) the CTAs get fatter instead of more numerous, so that
  // the per-CTA prologue (statistics + gamma / beta into shared memory, two barriers) is amortised over several
  // iterations of loads in flight (ncu, B = 32: 1.9 TB/s with 1024-vector CTAs)
  const long total_vec = static_cast<long>(B) * HW * (C / 8);
  const long per_cta = std::max<long>(1024, total_vec / (148 * 8));
  int ppa = ppc;
  while (ppa > 1 && static_cast<long>(ppa) * (C / 8) > per_cta && HW % (ppa / 2) == 0) ppa >>= 1;
  // block = a multiple of both the warp and the vectors per pixel (every thread keeps its channel vector)
  static const int apply_mode = [] { const char* e = getenv("PNP_GN_APPLY"); return e ? atoi(e) : 3; }();  // 0 old, 1 vec, 3 vec + reverse
  int lcm = nvec;
  while (lcm % 32) lcm += nvec;
  if (apply_mode != 0 && lcm <= 512) {
    const int threads_a = lcm * ((256 + lcm - 1) / lcm);
    PNP_CUDA(launch_k(gn_apply_vec_kernel<4>, dim3(HW / ppa, B), dim3(threads_a), 0, s, x0, C0, x1, C1, HW, mean_rstd, gamma,
                      beta, do_silu ? 1 : 0, out, ppa, (apply_mode & 2) ? 1 : 0));
    return 0;
  }
  PNP_CUDA(launch_k(gn_apply_kernel, dim3(HW / ppa, B), dim3(256), sm2, s, x0, C0, x1, C1, HW, mean_rstd, gamma, beta,
                    do_silu ? 1 : 0, out, ppa));
  return 0;
}

size_t groupnorm_workspace_floats(int B, int HW) {
  return groupnorm_partials_floats(B, HW) + 64 + 64 * GN_GROUPS * 2;
}

int layernorm_launch(const __half* x, int rows, int C, const float* gamma, const float* beta, float eps, __half* out,
                     cudaStream_t s) {
  PNP_CHECK(C % 8 == 0 && C <= 8 * 32 * 5, "layernorm: C");
  const int threads = 256;
  const int vpl = (C / 8 + 31) / 32;
  const int R = vpl <= 2 ? 4 : 2;
  const int warps = (rows + R - 1) / R;
  const int blocks = std::max(1, std::min((warps * 32 + threads - 1) / threads, 148 * 8));
  switch (vpl) {
    case 1: PNP_CUDA(launch_k(ln_kernel<1, 4>, dim3(blocks), dim3(threads), 0, s, x, rows, C, gamma, beta, eps, out)); break;
    case 2: PNP_CUDA(launch_k(ln_kernel<2, 4>, dim3(blocks), dim3(threads), 0, s, x, rows, C, gamma, beta, eps, out)); break;
    case 3: PNP_CUDA(launch_k(ln_kernel<3, 2>, dim3(blocks), dim3(threads), 0, s, x, rows, C, gamma, beta, eps, out)); break;
    case 4: PNP_CUDA(launch_k(ln_kernel<4, 2>, dim3(blocks), dim3(threads), 0, s, x, rows, C, gamma, beta, eps, out)); break;
    default: PNP_CUDA(launch_k(ln_kernel<5, 2>, dim3(blocks), dim3(threads), 0, s, x, rows, C, gamma, beta, eps, out)); break;
  }
  PNP_CUDA(cudaGetLastError());
  return 0;
}

int upsample2x_launch(const __half* x, int B, int H, int W, int C, __half* out, cudaStream_t s) {
  PNP_CHECK(C % 8 == 0, "upsample: C");
  const size_t total = static_cast<size_t>(B) * 4 * H * W * (C / 8);
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 16));
  PNP_CUDA(launch_k(upsample2x_kernel, dim3(blocks), dim3(256), 0, s, reinterpret_cast<const uint4*>(x), B, H, W, C / 8,
                    reinterpret_cast<uint4*>(out)));
  return 0;
}

int im2col_s2_launch(const __half* x, int B, int H, int W, int C, __half* out, cudaStream_t s) {
  PNP_CHECK(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "im2col: shape");
  const size_t total = static_cast<size_t>(B) * (H / 2) * (W / 2) * 9 * (C / 8);
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 16));
  PNP_CUDA(launch_k(im2col_s2_kernel, dim3(blocks), dim3(256), 0, s, reinterpret_cast<const uint4*>(x), B, H, W, C / 8,
                    reinterpret_cast<uint4*>(out)));
  return 0;
}

int conv_in_launch(const float* x_nchw, int B, int H, int W, const float* w, const float* bias, __half* out,
                   cudaStream_t s) {
  const int threads = 240;  // 40 channel-octets x 6 groups of 8 pixels
  PNP_CHECK(W % CIN_PX == 0, "conv_in: W must be a multiple of 8");
  const size_t ngroups = static_cast<size_t>(B) * H * W / CIN_PX;
  const int blocks = static_cast<int>(std::min<size_t>((ngroups + 5) / 6, 148 * 2));
  PNP_CUDA(launch_k(conv_in_kernel, dim3(blocks), dim3(threads), 0, s, x_nchw, B, H, W, w, bias, out));
  return 0;
}

}  // namespace pnp
