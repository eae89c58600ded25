// SD-1.x AutoencoderKL (the VAE the editors call around the loop: utils/utils.py:58-80 `image2latent` / `latent2image`,
// one encode + five decodes per edited image) on the same building blocks as the UNet engine: every 3x3 convolution is
// the tcgen05 implicit GEMM of gemm_sm100.cu (128 / 256 / 512 channels; images up to 512 pixels wide: one tile = 128
// consecutive pixels of a row), GroupNorm(32, eps 1e-6)+SiLU and nearest 2x upsampling are the kernels of norm.cu, the
// single-head 512-channel mid-block attention is two GEMMs around a row softmax.  Arithmetic spec: the reference's
// vendored diffusers 0.3.0, models/edict/my_diffusers/models/vae.py:54-131 (Encoder), :133-210 (Decoder), :480-557
// (AutoencoderKL, quant / post_quant 1x1), unet_blocks.py (DownEncoderBlock2D / UpDecoderBlock2D / UNetMidBlock2D),
// resnet.py:64-97 (stride-2 conv with right/bottom padding), attention.py:9-93 (AttentionBlock).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/pnpinv.h"
#include "pnp_internal.h"

namespace pnp {
namespace {

__device__ __forceinline__ uint4 pack8h(const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
  return u;
}

// ------------------------------------------------------------------ 3x3 conv from a few fp32 NCHW planes to fp16 NHWC
// (encoder.conv_in 3 -> 128, decoder.conv_in 4 -> 512).  One thread = 8 output channels x 4 consecutive pixels; the
// [k][co] weights (k = ci*9 + tap) are staged in shared memory once per CTA.
constexpr int SC_PX = 4;
__global__ void __launch_bounds__(256) small_conv3_kernel(const float* __restrict__ x, int B, int Cin, int H, int W,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          int Cout, __half* __restrict__ out) {
  extern __shared__ __align__(16) float ws[];  // [Cin*9][Cout]
  const int K = Cin * 9;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int nvec = Cout / 8;
  const int gpr = W / SC_PX;
  const size_t nitems = static_cast<size_t>(B) * H * gpr * nvec;
  for (size_t it = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; it < nitems;
       it += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(it % nvec);
    size_t g = it / nvec;
    const int x0 = static_cast<int>(g % gpr) * SC_PX;
    g /= gpr;
    const int yy = static_cast<int>(g % H);
    const int b = static_cast<int>(g / H);
    float acc[SC_PX][8];
#pragma unroll
    for (int j = 0; j < SC_PX; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[j][e] = bias[v * 8 + e];
    for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yi = yy + dy - 1;
        float xin[SC_PX + 2];
        const float* row = x + ((static_cast<size_t>(b) * Cin + ci) * H + yi) * W;
#pragma unroll
        for (int j = 0; j < SC_PX + 2; ++j) {
          const int xi = x0 + j - 1;
          xin[j] = (yi >= 0 && yi < H && xi >= 0 && xi < W) ? __ldg(row + xi) : 0.f;
        }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float* wp = ws + (ci * 9 + dy * 3 + dx) * Cout + v * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wp);
          const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
          for (int j = 0; j < SC_PX; ++j) {
            const float a = xin[j + dx];
            acc[j][0] += a * w0.x; acc[j][1] += a * w0.y; acc[j][2] += a * w0.z; acc[j][3] += a * w0.w;
            acc[j][4] += a * w1.x; acc[j][5] += a * w1.y; acc[j][6] += a * w1.z; acc[j][7] += a * w1.w;
          }
        }
      }
    }
    const size_t pix0 = (static_cast<size_t>(b) * H + yy) * W + x0;
#pragma unroll
    for (int j = 0; j < SC_PX; ++j) *reinterpret_cast<uint4*>(out + (pix0 + j) * Cout + v * 8) = pack8h(acc[j]);
  }
}

// 1x1 conv over <= 8 fp32 NCHW planes (quant_conv 8 -> 8, post_quant_conv 4 -> 4): out[b][co][p] = bias + sum w[co][ci] x
__global__ void pointwise_f32_kernel(const float* __restrict__ x, int B, int Cin, int Cout, int HW,
                                     const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out) {
  const size_t total = static_cast<size_t>(B) * HW;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / HW, p = i - b * HW;
    float xi[8];
    for (int ci = 0; ci < Cin; ++ci) xi[ci] = x[(b * Cin + ci) * HW + p];
    for (int co = 0; co < Cout; ++co) {
      float a = bias[co];
      for (int ci = 0; ci < Cin; ++ci) a += w[co * Cin + ci] * xi[ci];
      out[(b * Cout + co) * HW + p] = a;
    }
  }
}

// in-place row softmax of an fp16 [rows, cols] matrix (cols % 8 == 0, cols <= 256 * VPL): one warp per row, the row lives
// in registers between the two passes (AttentionBlock: softmax over all keys, attention.py:77-79)
template <int VPL>  // 16-byte vectors per lane
__global__ void __launch_bounds__(256) row_softmax_kernel(__half* __restrict__ s, int rows, int cols) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  uint4* row = reinterpret_cast<uint4*>(s + static_cast<size_t>(warp) * cols);
  float v[VPL][8];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const bool in = (i * 32 + lane) * 8 < cols;
    uint4 u = make_uint4(0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u, 0xfc00fc00u);  // -inf halves: exp -> 0
    if (in) u = row[i * 32 + lane];
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h[e]);
      v[i][2 * e] = f.x;
      v[i][2 * e + 1] = f.y;
      mx = fmaxf(mx, fmaxf(f.x, f.y));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[i][e] = __expf(v[i][e] - mx);
      sum += v[i][e];
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[i][e] *= inv;
    if ((i * 32 + lane) * 8 < cols) row[i * 32 + lane] = pack8h(v[i]);
  }
}

// out[c][r] = in[r][c]  (fp16; V -> V^T so that P.V is a K-major GEMM), 32x32 tiles through shared memory
__global__ void transpose_f16_kernel(const __half* __restrict__ in, int R, int Cc, int ld_in, __half* __restrict__ out,
                                     int ld_out) {
  __shared__ __half t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const size_t zin = static_cast<size_t>(blockIdx.z) * R * ld_in, zout = static_cast<size_t>(blockIdx.z) * Cc * ld_out;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < R && c < Cc) ? in[zin + static_cast<size_t>(r) * ld_in + c] : __float2half(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < R) out[zout + static_cast<size_t>(c) * ld_out + r] = t[threadIdx.x][i];
  }
}

// stride-2 3x3 conv input gather with padding 0 on the left/top and 1 on the right/bottom (Downsample2D with padding=0:
// F.pad(x, (0,1,0,1)) then conv stride 2, resnet.py:88-97): out[(b,yo,xo)][tap*C + c] = x[b, 2yo+dy, 2xo+dx, c]
__global__ void im2col_s2_pad01_kernel(const uint4* __restrict__ x, int B, int H, int W, int nvec, uint4* __restrict__ out) {
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
    const int yi = 2 * yo + tap / 3, xi = 2 * xo + tap % 3;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (yi < H && xi < W) u = x[((static_cast<size_t>(b) * H + yi) * W + xi) * nvec + v];
    out[idx] = u;
  }
}

// ------------------------------------------------------------------ host side
struct VResnet {
  int cin, cout;
  float *g1, *b1, *g2, *b2;
  __half *w1, *w2;  // [cout, 9*cin], [cout, 9*cout (+cin shortcut)]
  float *bias1, *bias2;
};
struct VAttn {
  float *gn_g, *gn_b;
  __half *wq, *wk, *wv, *wo;
  float *bq, *bk, *bv, *bo;
};
struct VPlan {
  std::vector<std::function<int(cudaStream_t)>> ops;
  std::vector<std::unique_ptr<GemmPlan>> gemms;
  std::vector<void*> bufs;
  float* in32 = nullptr;   // [B, 3 or 4, H, W]
  float* out32 = nullptr;  // decode: [B,3,8h,8w]; encode: [B,8,h,w] moments after quant_conv
  float* mid32 = nullptr;  // decode: post_quant output; encode: conv_out planes before quant_conv
  float* gemm_ws = nullptr;
  int* gemm_counters = nullptr;
  int launches = 0;
  double flops = 0.0;
};

}  // namespace
}  // namespace pnp

using namespace pnp;

struct pnp_vae {
  int device = 0, num_sms = 148;
  bool finalized = false;
  std::unordered_map<std::string, std::vector<__half>> host;
  std::vector<void*> allocs;
  // encoder
  float *e_in_w = nullptr, *e_in_b = nullptr;
  std::vector<VResnet> e_res;  // 8 down + 2 mid
  __half* e_down_w[3] = {nullptr, nullptr, nullptr};
  float* e_down_b[3] = {nullptr, nullptr, nullptr};
  VAttn e_attn;
  float *e_no_g = nullptr, *e_no_b = nullptr, *e_out_b = nullptr;
  __half* e_out_w = nullptr;
  float *q_w = nullptr, *q_b = nullptr, *pq_w = nullptr, *pq_b = nullptr;
  // decoder
  float *d_in_w = nullptr, *d_in_b = nullptr;
  std::vector<VResnet> d_res;  // 2 mid + 12 up
  __half* d_up_w[3] = {nullptr, nullptr, nullptr};
  float* d_up_b[3] = {nullptr, nullptr, nullptr};
  VAttn d_attn;
  float *d_no_g = nullptr, *d_no_b = nullptr, *d_out_b = nullptr;
  __half* d_out_w = nullptr;
  float* gn_ws = nullptr;
  std::map<std::array<int, 4>, std::unique_ptr<VPlan>> plans;  // (encode?, B, H, W)
  int64_t launches = 0;

  template <typename T>
  T* dalloc(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
    allocs.push_back(p);
    return static_cast<T*>(p);
  }
};

namespace pnp {
namespace {

const int kVaeOut[4] = {128, 256, 512, 512};

const std::vector<__half>* vfind(pnp_vae* e, const std::string& n) {
  auto it = e->host.find(n);
  return it == e->host.end() ? nullptr : &it->second;
}
std::vector<float> vf32(const std::vector<__half>& v) {
  std::vector<float> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __half2float(v[i]);
  return o;
}
float* vup32(pnp_vae* e, const std::vector<float>& v) {
  float* d = e->dalloc<float>(v.size());
  if (d && cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  return d;
}
__half* vup16(pnp_vae* e, const std::vector<__half>& v) {
  __half* d = e->dalloc<__half>(v.size());
  if (d && cudaMemcpy(d, v.data(), v.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  return d;
}
// the named tensor with `numel` values, or null + error
const std::vector<__half>* need(pnp_vae* e, const std::string& n, int64_t numel, bool* ok) {
  const auto* p = vfind(e, n);
  if (p == nullptr || static_cast<int64_t>(p->size()) != numel) {
    set_last_error("vae: parameter missing or wrong size: " + n);
    *ok = false;
    return nullptr;
  }
  return p;
}
// OIHW 3x3 -> [co][(ky*3+kx)*cin + ci] (+ 1x1 shortcut columns), the layout of the implicit-GEMM conv
std::vector<__half> vpack3(const std::vector<__half>& w, int cout, int cin, const std::vector<__half>* sc, int sc_cin) {
  const size_t K = static_cast<size_t>(9) * cin + (sc ? sc_cin : 0);
  std::vector<__half> p(static_cast<size_t>(cout) * K);
  for (int co = 0; co < cout; ++co) {
    __half* row = p.data() + static_cast<size_t>(co) * K;
    const __half* src = w.data() + static_cast<size_t>(co) * cin * 9;
    for (int ci = 0; ci < cin; ++ci)
      for (int tap = 0; tap < 9; ++tap) row[static_cast<size_t>(tap) * cin + ci] = src[ci * 9 + tap];
    if (sc) memcpy(row + static_cast<size_t>(9) * cin, sc->data() + static_cast<size_t>(co) * sc_cin, sc_cin * sizeof(__half));
  }
  return p;
}
// small conv weights [co][ci][tap] -> fp32 [k = ci*9+tap][co]
std::vector<float> vpack_small(const std::vector<__half>& w, int cout, int cin) {
  std::vector<float> p(w.size());
  for (int co = 0; co < cout; ++co)
    for (int k = 0; k < cin * 9; ++k) p[static_cast<size_t>(k) * cout + co] = __half2float(w[static_cast<size_t>(co) * cin * 9 + k]);
  return p;
}

bool load_resnet(pnp_vae* e, const std::string& n, int cin, int cout, VResnet* r) {
  bool ok = true;
  r->cin = cin;
  r->cout = cout;
  const auto* g1 = need(e, n + ".norm1.weight", cin, &ok);
  const auto* b1 = need(e, n + ".norm1.bias", cin, &ok);
  const auto* w1 = need(e, n + ".conv1.weight", static_cast<int64_t>(cout) * cin * 9, &ok);
  const auto* c1 = need(e, n + ".conv1.bias", cout, &ok);
  const auto* g2 = need(e, n + ".norm2.weight", cout, &ok);
  const auto* b2 = need(e, n + ".norm2.bias", cout, &ok);
  const auto* w2 = need(e, n + ".conv2.weight", static_cast<int64_t>(cout) * cout * 9, &ok);
  const auto* c2 = need(e, n + ".conv2.bias", cout, &ok);
  const std::vector<__half>* sw = nullptr;
  const std::vector<__half>* sb = nullptr;
  if (cin != cout) {
    sw = need(e, n + ".conv_shortcut.weight", static_cast<int64_t>(cout) * cin, &ok);
    sb = need(e, n + ".conv_shortcut.bias", cout, &ok);
  }
  if (!ok) return false;
  r->g1 = vup32(e, vf32(*g1)); r->b1 = vup32(e, vf32(*b1));
  r->g2 = vup32(e, vf32(*g2)); r->b2 = vup32(e, vf32(*b2));
  r->w1 = vup16(e, vpack3(*w1, cout, cin, nullptr, 0));
  r->bias1 = vup32(e, vf32(*c1));
  std::vector<float> bias2 = vf32(*c2);
  if (sw) {
    r->w2 = vup16(e, vpack3(*w2, cout, cout, sw, cin));
    const std::vector<float> sbf = vf32(*sb);
    for (int i = 0; i < cout; ++i) bias2[i] += sbf[i];
  } else {
    r->w2 = vup16(e, vpack3(*w2, cout, cout, nullptr, 0));
  }
  r->bias2 = vup32(e, bias2);
  if (!(r->g1 && r->b1 && r->g2 && r->b2 && r->w1 && r->w2 && r->bias1 && r->bias2)) {
    set_last_error("vae: upload failed (out of memory?)");
    return false;
  }
  return true;
}

bool load_attn(pnp_vae* e, const std::string& n, int c, VAttn* a) {
  bool ok = true;
  const auto* g = need(e, n + ".group_norm.weight", c, &ok);
  const auto* b = need(e, n + ".group_norm.bias", c, &ok);
  const char* names[4] = {"query", "key", "value", "proj_attn"};
  const std::vector<__half>* w[4];
  const std::vector<__half>* bb[4];
  for (int i = 0; i < 4; ++i) {
    w[i] = need(e, n + "." + names[i] + ".weight", static_cast<int64_t>(c) * c, &ok);
    bb[i] = need(e, n + "." + names[i] + ".bias", c, &ok);
  }
  if (!ok) return false;
  a->gn_g = vup32(e, vf32(*g)); a->gn_b = vup32(e, vf32(*b));
  a->wq = vup16(e, *w[0]); a->wk = vup16(e, *w[1]); a->wv = vup16(e, *w[2]); a->wo = vup16(e, *w[3]);
  a->bq = vup32(e, vf32(*bb[0])); a->bk = vup32(e, vf32(*bb[1])); a->bv = vup32(e, vf32(*bb[2])); a->bo = vup32(e, vf32(*bb[3]));
  return a->gn_g && a->gn_b && a->wq && a->wk && a->wv && a->wo && a->bq && a->bk && a->bv && a->bo;
}

// conv_out as a zero-padded 64-column tile with fp32 NCHW planes (like the UNet's conv_out)
bool load_conv_out(pnp_vae* e, const std::string& n, int cout, int cin, __half** w_out, float** b_out) {
  bool ok = true;
  const auto* w = need(e, n + ".weight", static_cast<int64_t>(cout) * cin * 9, &ok);
  const auto* b = need(e, n + ".bias", cout, &ok);
  if (!ok) return false;
  std::vector<__half> p = vpack3(*w, cout, cin, nullptr, 0);
  p.resize(static_cast<size_t>(64) * 9 * cin, __float2half(0.f));
  std::vector<float> bf = vf32(*b);
  bf.resize(64, 0.f);
  *w_out = vup16(e, p);
  *b_out = vup32(e, bf);
  return *w_out && *b_out;
}

int vae_finalize(pnp_vae* e) {
  bool ok = true;
  // ---- encoder
  {
    const auto* w = need(e, "encoder.conv_in.weight", 128 * 3 * 9, &ok);
    const auto* b = need(e, "encoder.conv_in.bias", 128, &ok);
    if (!ok) return -2;
    e->e_in_w = vup32(e, vpack_small(*w, 128, 3));
    e->e_in_b = vup32(e, vf32(*b));
  }
  int cin = 128;
  for (int i = 0; i < 4; ++i) {
    const int cout = kVaeOut[i];
    for (int j = 0; j < 2; ++j) {
      VResnet r;
      if (!load_resnet(e, "encoder.down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), j == 0 ? cin : cout, cout, &r))
        return -2;
      e->e_res.push_back(r);
    }
    if (i < 3) {
      const std::string d = "encoder.down_blocks." + std::to_string(i) + ".downsamplers.0.conv";
      const auto* w = need(e, d + ".weight", static_cast<int64_t>(cout) * cout * 9, &ok);
      const auto* b = need(e, d + ".bias", cout, &ok);
      if (!ok) return -2;
      e->e_down_w[i] = vup16(e, vpack3(*w, cout, cout, nullptr, 0));
      e->e_down_b[i] = vup32(e, vf32(*b));
    }
    cin = cout;
  }
  for (int j = 0; j < 2; ++j) {
    VResnet r;
    if (!load_resnet(e, "encoder.mid_block.resnets." + std::to_string(j), 512, 512, &r)) return -2;
    e->e_res.push_back(r);
  }
  if (!load_attn(e, "encoder.mid_block.attentions.0", 512, &e->e_attn)) return -2;
  {
    const auto* g = need(e, "encoder.conv_norm_out.weight", 512, &ok);
    const auto* b = need(e, "encoder.conv_norm_out.bias", 512, &ok);
    if (!ok) return -2;
    e->e_no_g = vup32(e, vf32(*g));
    e->e_no_b = vup32(e, vf32(*b));
    if (!load_conv_out(e, "encoder.conv_out", 8, 512, &e->e_out_w, &e->e_out_b)) return -2;
    const auto* qw = need(e, "quant_conv.weight", 64, &ok);
    const auto* qb = need(e, "quant_conv.bias", 8, &ok);
    const auto* pw = need(e, "post_quant_conv.weight", 16, &ok);
    const auto* pb = need(e, "post_quant_conv.bias", 4, &ok);
    if (!ok) return -2;
    e->q_w = vup32(e, vf32(*qw)); e->q_b = vup32(e, vf32(*qb));
    e->pq_w = vup32(e, vf32(*pw)); e->pq_b = vup32(e, vf32(*pb));
  }
  // ---- decoder
  {
    const auto* w = need(e, "decoder.conv_in.weight", 512 * 4 * 9, &ok);
    const auto* b = need(e, "decoder.conv_in.bias", 512, &ok);
    if (!ok) return -2;
    e->d_in_w = vup32(e, vpack_small(*w, 512, 4));
    e->d_in_b = vup32(e, vf32(*b));
  }
  for (int j = 0; j < 2; ++j) {
    VResnet r;
    if (!load_resnet(e, "decoder.mid_block.resnets." + std::to_string(j), 512, 512, &r)) return -2;
    e->d_res.push_back(r);
  }
  if (!load_attn(e, "decoder.mid_block.attentions.0", 512, &e->d_attn)) return -2;
  const int rev[4] = {512, 512, 256, 128};
  cin = 512;
  for (int i = 0; i < 4; ++i) {
    const int cout = rev[i];
    for (int j = 0; j < 3; ++j) {
      VResnet r;
      if (!load_resnet(e, "decoder.up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), j == 0 ? cin : cout, cout, &r))
        return -2;
      e->d_res.push_back(r);
    }
    if (i < 3) {
      const std::string u = "decoder.up_blocks." + std::to_string(i) + ".upsamplers.0.conv";
      const auto* w = need(e, u + ".weight", static_cast<int64_t>(cout) * cout * 9, &ok);
      const auto* b = need(e, u + ".bias", cout, &ok);
      if (!ok) return -2;
      e->d_up_w[i] = vup16(e, vpack3(*w, cout, cout, nullptr, 0));
      e->d_up_b[i] = vup32(e, vf32(*b));
    }
    cin = cout;
  }
  {
    const auto* g = need(e, "decoder.conv_norm_out.weight", 128, &ok);
    const auto* b = need(e, "decoder.conv_norm_out.bias", 128, &ok);
    if (!ok) return -2;
    e->d_no_g = vup32(e, vf32(*g));
    e->d_no_b = vup32(e, vf32(*b));
    if (!load_conv_out(e, "decoder.conv_out", 3, 128, &e->d_out_w, &e->d_out_b)) return -2;
  }
  e->host.clear();
  const size_t wsf = groupnorm_workspace_floats(PNP_MAX_BATCH, 4096) + 64 * 1024;
  e->gn_ws = e->dalloc<float>(wsf);
  PNP_CHECK(e->gn_ws != nullptr, "vae: alloc failed");
  PNP_CUDA(cudaMemset(e->gn_ws, 0, wsf * sizeof(float)));
  e->finalized = true;
  return 0;
}

// ------------------------------------------------------------------ plan builder
struct VBuilder {
  pnp_vae* e;
  VPlan* pl;
  int B;
  int rc = 0;

  __half* buf(size_t elems) {
    void* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(elems, 64) * sizeof(__half)) != cudaSuccess) {
      rc = -1;
      set_last_error("vae plan: cudaMalloc failed");
      return nullptr;
    }
    pl->bufs.push_back(p);
    return static_cast<__half*>(p);
  }
  float* buf32(size_t elems) { return reinterpret_cast<float*>(buf(2 * elems)); }
  void op(int kernels, std::function<int(cudaStream_t)> f) {
    pl->ops.push_back(std::move(f));
    pl->launches += kernels;
  }
  void gemm(const ASource* srcs, int nsrc, int taps, bool linear, int b, int h, int w, const __half* wt, int n, int ktot,
            const GemmEpilogue& ep) {
    if (rc) return;
    auto gp = std::make_unique<GemmPlan>();
    rc = gemm_plan_create(gp.get(), srcs, nsrc, taps, linear, b, h, w, wt, n, ktot, ep, 0, e->num_sms, 0);
    if (rc) return;
    GemmPlan* raw = gp.get();
    pl->gemms.push_back(std::move(gp));
    pl->flops += 2.0 * raw->p.M * static_cast<double>(n) * ktot;
    op(1, [raw](cudaStream_t s) { return gemm_launch(*raw, s); });
  }
  void conv3(const __half* x, int cin, int h, int w, const __half* wt, const float* bias, int cout, __half* out,
             const __half* residual = nullptr, const __half* sc_src = nullptr, int sc_c = 0) {
    GemmEpilogue ep;
    ep.bias = bias;
    ep.out = out;
    ep.ldc = cout;
    ep.residual = residual;
    ep.ldr = cout;
    ASource s[2];
    s[0] = ASource{x, cin, cin};
    int ns = 1, ktot = 9 * cin;
    if (sc_src) {
      s[ns++] = ASource{sc_src, sc_c, sc_c};
      ktot += sc_c;
    }
    gemm(s, ns, 9, false, B, h, w, wt, cout, ktot, ep);
  }
  void groupnorm(const __half* x, int c, int hw, const float* g, const float* b, bool silu, __half* out) {
    pnp_vae* en = e;
    const int Bn = B;
    op(groupnorm_kernel_count(c, Bn, hw),
       [=](cudaStream_t s) { return groupnorm_launch(x, c, nullptr, 0, Bn, hw, g, b, 1e-6f, silu, out, en->gn_ws, s); });
  }
  // resnet.py:331-365 without time embedding, GroupNorm eps 1e-6
  void resnet(const VResnet& r, const __half* x, int h, int w, __half* nrm, __half* h1, __half* out) {
    groupnorm(x, r.cin, h * w, r.g1, r.b1, true, nrm);
    conv3(nrm, r.cin, h, w, r.w1, r.bias1, r.cout, h1);
    groupnorm(h1, r.cout, h * w, r.g2, r.b2, true, nrm);
    if (r.cin != r.cout)
      conv3(nrm, r.cout, h, w, r.w2, r.bias2, r.cout, out, nullptr, x, r.cin);
    else
      conv3(nrm, r.cout, h, w, r.w2, r.bias2, r.cout, out, x);
  }
  // attention.py:9-93: GroupNorm -> q,k,v linear -> softmax(q k^T / sqrt(C)) v -> proj + residual, ONE head of 512 channels
  void attention(const VAttn& a, const __half* x, int n_tok, __half* nrm, __half* q, __half* k, __half* v, __half* vt,
                 __half* sc, __half* o, __half* out) {
    const int c = 512, M = B * n_tok;
    groupnorm(x, c, n_tok, a.gn_g, a.gn_b, false, nrm);
    auto lin = [&](const __half* in, const __half* wt, const float* bias, __half* dst, const __half* res) {
      GemmEpilogue ep;
      ep.bias = bias; ep.out = dst; ep.ldc = c; ep.residual = res; ep.ldr = c;
      ASource s{in, c, c};
      gemm(&s, 1, 1, true, 1, 1, M, wt, c, c, ep);
    };
    lin(nrm, a.wq, a.bq, q, nullptr);
    lin(nrm, a.wk, a.bk, k, nullptr);
    lin(nrm, a.wv, a.bv, v, nullptr);
    const int Bn = B;
    op(1, [=](cudaStream_t s) {
      transpose_f16_kernel<<<dim3((c + 31) / 32, (n_tok + 31) / 32, Bn), dim3(32, 8), 0, s>>>(v, n_tok, c, c, vt, n_tok);
      PNP_CUDA(cudaGetLastError());
      return 0;
    });
    for (int b = 0; b < B; ++b) {  // per image: S = Q K^T / sqrt(C) ; P = softmax(S) ; O = P V
      const __half* qb = q + static_cast<size_t>(b) * n_tok * c;
      const __half* kb = k + static_cast<size_t>(b) * n_tok * c;
      const __half* vtb = vt + static_cast<size_t>(b) * c * n_tok;
      __half* ob = o + static_cast<size_t>(b) * n_tok * c;
      {
        GemmEpilogue ep;
        ep.out = sc; ep.ldc = n_tok; ep.out_scale = 1.0f / sqrtf(static_cast<float>(c));
        ASource s{qb, c, c};
        gemm(&s, 1, 1, true, 1, 1, n_tok, kb, n_tok, c, ep);
      }
      op(1, [=](cudaStream_t s) {
        const int warps_per_block = 8;
        const int blocks = (n_tok + warps_per_block - 1) / warps_per_block;
        if (n_tok <= 256) row_softmax_kernel<1><<<blocks, 256, 0, s>>>(sc, n_tok, n_tok);
        else if (n_tok <= 1024) row_softmax_kernel<4><<<blocks, 256, 0, s>>>(sc, n_tok, n_tok);
        else if (n_tok <= 4096) row_softmax_kernel<16><<<blocks, 256, 0, s>>>(sc, n_tok, n_tok);
        else { set_last_error("vae attention: at most 4096 tokens"); return -2; }
        PNP_CUDA(cudaGetLastError());
        return 0;
      });
      {
        GemmEpilogue ep;
        ep.out = ob; ep.ldc = c;
        ASource s{sc, n_tok, n_tok};
        gemm(&s, 1, 1, true, 1, 1, n_tok, vtb, c, n_tok, ep);
      }
    }
    lin(o, a.wo, a.bo, out, x);
  }
};

int finish_plan(pnp_vae* e, VBuilder& vb) {
  if (vb.rc) return vb.rc;
  size_t ws = 0;
  for (auto& g : vb.pl->gemms) ws = std::max(ws, gemm_ws_floats(*g));
  if (ws > 0) {
    PNP_CUDA(cudaMalloc(reinterpret_cast<void**>(&vb.pl->gemm_ws), ws * sizeof(float)));
    vb.pl->bufs.push_back(vb.pl->gemm_ws);
    PNP_CUDA(cudaMalloc(reinterpret_cast<void**>(&vb.pl->gemm_counters), kGemmMaxCounters * sizeof(int)));
    vb.pl->bufs.push_back(vb.pl->gemm_counters);
    PNP_CUDA(cudaMemset(vb.pl->gemm_counters, 0, kGemmMaxCounters * sizeof(int)));
    for (auto& g : vb.pl->gemms) gemm_set_workspace(g.get(), vb.pl->gemm_ws, vb.pl->gemm_counters);
  }
  (void)e;
  return 0;
}

int small_conv3_launch(const float* x, int B, int Cin, int H, int W, const float* w, const float* bias, int Cout,
                       __half* out, cudaStream_t s) {
  PNP_CHECK(W % SC_PX == 0 && Cout % 8 == 0 && Cin <= 4, "vae conv_in: shape");
  const size_t smem = static_cast<size_t>(Cin) * 9 * Cout * sizeof(float);
  static bool attr = false;
  if (!attr) {
    PNP_CUDA(cudaFuncSetAttribute(small_conv3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  PNP_CHECK(smem <= 96 * 1024, "vae conv_in: weights do not fit shared memory");
  const size_t items = static_cast<size_t>(B) * H * (W / SC_PX) * (Cout / 8);
  const int blocks = static_cast<int>(std::min<size_t>((items + 255) / 256, 148 * 8));
  small_conv3_kernel<<<blocks, 256, smem, s>>>(x, B, Cin, H, W, w, bias, Cout, out);
  PNP_CUDA(cudaGetLastError());
  return 0;
}

// decoder (vae.py:133-210): z [B,4,h,w] fp32 -> image [B,3,8h,8w] fp32
int build_decode(pnp_vae* e, int B, int h, int w, VPlan* pl) {
  VBuilder vb{e, pl, B};
  PNP_CHECK(h == w && (h == 8 || h == 16 || h == 32 || h == 64), "vae decode: latent must be 8x8, 16x16, 32x32 or 64x64");
  const size_t top = static_cast<size_t>(B) * 64 * h * w;  // pixels of the output image
  pl->in32 = vb.buf32(static_cast<size_t>(B) * 4 * h * w);
  pl->mid32 = vb.buf32(static_cast<size_t>(B) * 4 * h * w);
  pl->out32 = vb.buf32(static_cast<size_t>(B) * 3 * 64 * h * w);
  // activation arena: the largest tensors are 128 ch at 8h x 8w and 256 ch at 8h x 8w (up_blocks.3 input after upsampling)
  __half* XA = vb.buf(top * 256);
  __half* XB = vb.buf(top * 256);
  __half* NRM = vb.buf(top * 256);
  __half* H1 = vb.buf(top * 128 + static_cast<size_t>(B) * 16 * h * w * 256);
  const int ntok = h * w;
  __half* Q = vb.buf(static_cast<size_t>(B) * ntok * 512);
  __half* K = vb.buf(static_cast<size_t>(B) * ntok * 512);
  __half* V = vb.buf(static_cast<size_t>(B) * ntok * 512);
  __half* VT = vb.buf(static_cast<size_t>(B) * ntok * 512);
  __half* O = vb.buf(static_cast<size_t>(B) * ntok * 512);
  __half* SC = vb.buf(static_cast<size_t>(ntok) * ntok);
  if (vb.rc) return vb.rc;
  const int HW = h * w;
  {
    const float* in = pl->in32; float* mid = pl->mid32;
    const float* pw = e->pq_w; const float* pb = e->pq_b;
    vb.op(1, [=](cudaStream_t s) {
      pointwise_f32_kernel<<<std::min((B * HW + 255) / 256, 1184), 256, 0, s>>>(in, B, 4, 4, HW, pw, pb, mid);
      PNP_CUDA(cudaGetLastError());
      return 0;
    });
    const float* wi = e->d_in_w; const float* bi = e->d_in_b;
    vb.op(1, [=](cudaStream_t s) { return small_conv3_launch(mid, B, 4, h, w, wi, bi, 512, XA, s); });
  }
  __half* x = XA;
  __half* y = XB;
  int ri = 0;
  vb.resnet(e->d_res[ri++], x, h, w, NRM, H1, y); std::swap(x, y);
  vb.attention(e->d_attn, x, ntok, NRM, Q, K, V, VT, SC, O, y); std::swap(x, y);
  vb.resnet(e->d_res[ri++], x, h, w, NRM, H1, y); std::swap(x, y);
  const int rev[4] = {512, 512, 256, 128};
  int ch = h, cw = w;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 3; ++j) {
      vb.resnet(e->d_res[ri++], x, ch, cw, NRM, H1, y);
      std::swap(x, y);
    }
    if (i < 3) {
      const __half* src = x; const int c = rev[i], hh = ch, ww = cw;
      __half* ups = NRM;  // the normalised-activation scratch is free between resnets
      vb.op(1, [=](cudaStream_t s) { return upsample2x_launch(src, B, hh, ww, c, ups, s); });
      ch *= 2; cw *= 2;
      vb.conv3(ups, c, ch, cw, e->d_up_w[i], e->d_up_b[i], c, y);
      std::swap(x, y);
    }
  }
  vb.groupnorm(x, 128, ch * cw, e->d_no_g, e->d_no_b, true, NRM);
  {
    GemmEpilogue ep;
    ep.bias = e->d_out_b;
    ep.out_f32_nchw4 = pl->out32;
    ep.out32_channels = 3;
    ASource s{NRM, 128, 128};
    vb.gemm(&s, 1, 9, false, B, ch, cw, e->d_out_w, 64, 9 * 128, ep);
  }
  return finish_plan(e, vb);
}

// encoder (vae.py:54-131) + quant_conv: image [B,3,H,W] fp32 in [-1,1] -> moments [B,8,H/8,W/8] (mean | logvar)
int build_encode(pnp_vae* e, int B, int H, int W, VPlan* pl) {
  VBuilder vb{e, pl, B};
  const int h = H / 8, w = W / 8;
  PNP_CHECK(H == W && (H == 64 || H == 128 || H == 256 || H == 512), "vae encode: image must be 64, 128, 256 or 512 pixels square");
  const size_t top = static_cast<size_t>(B) * H * W;
  pl->in32 = vb.buf32(top * 3);
  pl->mid32 = vb.buf32(static_cast<size_t>(B) * 8 * h * w);
  pl->out32 = vb.buf32(static_cast<size_t>(B) * 8 * h * w);
  __half* XA = vb.buf(top * 128);
  __half* XB = vb.buf(top * 128);
  __half* NRM = vb.buf(top * 128);
  __half* H1 = vb.buf(top * 128);
  __half* IM2 = vb.buf(top / 4 * 9 * 128);
  const int ntok = h * w;
  __half* Q = vb.buf(static_cast<size_t>(B) * ntok * 512);
  __half* K = vb.buf(static_cast<size_t>(B) * ntok * 512);
  __half* V = vb.buf(static_cast<size_t>(B) * ntok * 512);
  __half* VT = vb.buf(static_cast<size_t>(B) * ntok * 512);
  __half* O = vb.buf(static_cast<size_t>(B) * ntok * 512);
  __half* SC = vb.buf(static_cast<size_t>(ntok) * ntok);
  if (vb.rc) return vb.rc;
  {
    const float* in = pl->in32; const float* wi = e->e_in_w; const float* bi = e->e_in_b;
    vb.op(1, [=](cudaStream_t s) { return small_conv3_launch(in, B, 3, H, W, wi, bi, 128, XA, s); });
  }
  __half* x = XA;
  __half* y = XB;
  int ri = 0, ch = H, cw = W;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 2; ++j) {
      vb.resnet(e->e_res[ri++], x, ch, cw, NRM, H1, y);
      std::swap(x, y);
    }
    if (i < 3) {
      const __half* src = x; const int c = kVaeOut[i], hh = ch, ww = cw;
      vb.op(1, [=](cudaStream_t s) {
        const size_t total = static_cast<size_t>(B) * (hh / 2) * (ww / 2) * 9 * (c / 8);
        const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 16));
        im2col_s2_pad01_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const uint4*>(src), B, hh, ww, c / 8,
                                                      reinterpret_cast<uint4*>(IM2));
        PNP_CUDA(cudaGetLastError());
        return 0;
      });
      ch /= 2; cw /= 2;
      GemmEpilogue ep;
      ep.bias = e->e_down_b[i]; ep.out = y; ep.ldc = c;
      ASource s{IM2, 9 * c, 9 * c};
      vb.gemm(&s, 1, 1, true, 1, 1, B * ch * cw, e->e_down_w[i], c, 9 * c, ep);
      std::swap(x, y);
    }
  }
  vb.resnet(e->e_res[ri++], x, ch, cw, NRM, H1, y); std::swap(x, y);
  vb.attention(e->e_attn, x, ntok, NRM, Q, K, V, VT, SC, O, y); std::swap(x, y);
  vb.resnet(e->e_res[ri++], x, ch, cw, NRM, H1, y); std::swap(x, y);
  vb.groupnorm(x, 512, ch * cw, e->e_no_g, e->e_no_b, true, NRM);
  {
    GemmEpilogue ep;
    ep.bias = e->e_out_b;
    ep.out_f32_nchw4 = pl->mid32;
    ep.out32_channels = 8;
    ASource s{NRM, 512, 512};
    vb.gemm(&s, 1, 9, false, B, ch, cw, e->e_out_w, 64, 9 * 512, ep);
  }
  {
    const float* mid = pl->mid32; float* out = pl->out32;
    const float* qw = e->q_w; const float* qb = e->q_b;
    const int HW = ch * cw;
    vb.op(1, [=](cudaStream_t s) {
      pointwise_f32_kernel<<<std::min((B * HW + 255) / 256, 1184), 256, 0, s>>>(mid, B, 8, 8, HW, qw, qb, out);
      PNP_CUDA(cudaGetLastError());
      return 0;
    });
  }
  return finish_plan(e, vb);
}

int get_vplan(pnp_vae* e, bool encode, int B, int H, int W, VPlan** out) {
  const std::array<int, 4> key = {encode ? 1 : 0, B, H, W};
  auto it = e->plans.find(key);
  if (it == e->plans.end()) {
    auto pl = std::make_unique<VPlan>();
    int rc = encode ? build_encode(e, B, H, W, pl.get()) : build_decode(e, B, H, W, pl.get());
    if (rc) {
      for (void* p : pl->bufs) cudaFree(p);
      return rc;
    }
    it = e->plans.emplace(key, std::move(pl)).first;
  }
  *out = it->second.get();
  return 0;
}

}  // namespace
}  // namespace pnp

extern "C" {

int pnp_vae_create(int device_ordinal, pnp_vae** out) {
  PNP_CHECK(out != nullptr, "pnp_vae_create: out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e0 = cudaGetDeviceCount(&ndev);
  if (e0 != cudaSuccess || ndev == 0) {
    set_last_error(std::string("pnp_vae_create: no CUDA device available (") + cudaGetErrorString(e0) +
                   "); this library has no CPU fallback");
    return -1;
  }
  PNP_CHECK(device_ordinal >= 0 && device_ordinal < ndev, "pnp_vae_create: bad device ordinal");
  PNP_CUDA(cudaSetDevice(device_ordinal));
  cudaDeviceProp prop;
  PNP_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
  PNP_CHECK(prop.major == 10, "pnp_vae_create: this library is built for sm_100a (B200) only");
  auto* e = new pnp_vae();
  e->device = device_ordinal;
  e->num_sms = prop.multiProcessorCount;
  *out = e;
  return 0;
}

void pnp_vae_destroy(pnp_vae* h) {
  if (h == nullptr) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->plans)
    for (void* p : kv.second->bufs) cudaFree(p);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

int pnp_vae_load_param(pnp_vae* h, const char* name, const uint16_t* data_host, int64_t numel) {
  PNP_CHECK(h && name && data_host && numel > 0, "pnp_vae_load_param: bad argument");
  PNP_CHECK(!h->finalized, "pnp_vae_load_param: parameters already finalized");
  std::vector<__half> v(static_cast<size_t>(numel));
  memcpy(static_cast<void*>(v.data()), data_host, static_cast<size_t>(numel) * sizeof(uint16_t));
  h->host[name] = std::move(v);
  return 0;
}

int pnp_vae_finalize(pnp_vae* h) {
  PNP_CHECK(h != nullptr && !h->finalized, "pnp_vae_finalize: bad handle");
  PNP_CUDA(cudaSetDevice(h->device));
  return vae_finalize(h);
}

static int run_vplan(pnp_vae* h, VPlan* pl, cudaStream_t s) {
  for (auto& f : pl->ops) {
    int rc = f(s);
    if (rc) return rc;
  }
  h->launches += pl->launches;
  return 0;
}

int pnp_vae_decode(pnp_vae* h, const float* z_dev, int batch, int lat_h, int lat_w, float* image_out_dev, void* stream) {
  PNP_CHECK(h && h->finalized && z_dev && image_out_dev, "pnp_vae_decode: bad argument");
  PNP_CHECK(batch >= 1 && batch <= 8, "pnp_vae_decode: 1..8 images per call");
  VPlan* pl = nullptr;
  int rc = get_vplan(h, false, batch, lat_h, lat_w, &pl);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t zin = static_cast<size_t>(batch) * 4 * lat_h * lat_w * sizeof(float);
  PNP_CUDA(cudaMemcpyAsync(pl->in32, z_dev, zin, cudaMemcpyDeviceToDevice, s));
  rc = run_vplan(h, pl, s);
  if (rc) return rc;
  PNP_CUDA(cudaMemcpyAsync(image_out_dev, pl->out32, static_cast<size_t>(batch) * 3 * 64 * lat_h * lat_w * sizeof(float),
                           cudaMemcpyDeviceToDevice, s));
  return 0;
}

int pnp_vae_encode(pnp_vae* h, const float* image_dev, int batch, int img_h, int img_w, float* moments_out_dev,
                   void* stream) {
  PNP_CHECK(h && h->finalized && image_dev && moments_out_dev, "pnp_vae_encode: bad argument");
  PNP_CHECK(batch >= 1 && batch <= 8, "pnp_vae_encode: 1..8 images per call");
  VPlan* pl = nullptr;
  int rc = get_vplan(h, true, batch, img_h, img_w, &pl);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  PNP_CUDA(cudaMemcpyAsync(pl->in32, image_dev, static_cast<size_t>(batch) * 3 * img_h * img_w * sizeof(float),
                           cudaMemcpyDeviceToDevice, s));
  rc = run_vplan(h, pl, s);
  if (rc) return rc;
  PNP_CUDA(cudaMemcpyAsync(moments_out_dev, pl->out32,
                           static_cast<size_t>(batch) * 8 * (img_h / 8) * (img_w / 8) * sizeof(float),
                           cudaMemcpyDeviceToDevice, s));
  return 0;
}

int pnp_vae_kernel_launches(pnp_vae* h, int64_t* out) {
  PNP_CHECK(h && out, "null argument");
  *out = h->launches;
  return 0;
}

}  // extern "C"
