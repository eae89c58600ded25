// CLIP ViT-L/14 text encoder (the `model.text_encoder(input_ids)[0]` every loop of the reference calls before its first
// UNet step: models/p2p/inversion.py:42,50,296,304, models/p2p/p2p_guidance_forward.py:43,49,86,92,
// models/edict/edict_functions.py:818-838) on the building blocks of the UNet engine: the four projections of each of the
// 12 pre-LayerNorm blocks are the tcgen05 GEMM of gemm_sm100.cu (M = 77 tokens x prompts, bias / residual in the
// epilogue), LayerNorm is norm.cu's kernel; the 77-token causal attention (12 heads of 64) and the small element-wise
// steps are the kernels below.
//
// Arithmetic spec: `transformers` CLIPTextModel (a third-party dependency of the reference, pinned 4.19.2 in
// environment/edict_requirements.txt and unpinned for P2P / MasaCtrl; absent from /root/reference), configuration of the
// SD-1.x `text_encoder/`: hidden 768, 12 layers, 12 heads, MLP 3072 with quick_gelu, 77 positions, LayerNorm eps 1e-5,
// causal mask, final LayerNorm; output = last_hidden_state.  Restated for the CPU in oracle/clip_ref.py, which is pinned
// on the installed transformers implementation (tests/test_oracle_cpu.py, tests/golden/clip_text.npz).
#include <algorithm>
#include <array>
#include <cmath>
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

constexpr int CH = 768;      // hidden size
constexpr int CHEADS = 12;   // attention heads
constexpr int CHD = 64;      // head dim
constexpr int CFF = 3072;    // MLP width
constexpr int CTOK = 77;     // positions
constexpr int CKP = 80;      // padded key count of the transposed K tile

// x[row, :] = token_embedding[ids[row]] + position_embedding[row % 77]   (fp32 sum, fp16 storage)
__global__ void __launch_bounds__(96) clip_embed_kernel(const int* __restrict__ ids, const __half* __restrict__ tok,
                                                        const __half* __restrict__ pos, __half* __restrict__ x) {
  const int row = blockIdx.x;
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(tok + static_cast<size_t>(ids[row]) * CH) + threadIdx.x);
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(pos + static_cast<size_t>(row % CTOK) * CH) + threadIdx.x);
  const __half2* ha = reinterpret_cast<const __half2*>(&a);
  const __half2* hb = reinterpret_cast<const __half2*>(&b);
  uint4 o;
  __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 fa = __half22float2(ha[e]), fb = __half22float2(hb[e]);
    ho[e] = __floats2half2_rn(fa.x + fb.x, fa.y + fb.y);
  }
  reinterpret_cast<uint4*>(x + static_cast<size_t>(row) * CH)[threadIdx.x] = o;
}

// Causal self-attention over the 77 tokens of one (prompt, head): softmax(q k^T + mask) v, q pre-scaled by 1/sqrt(64)
// through the packed projection.  K^T and V of the head live in shared memory as fp32; one warp per query row, lanes
// over keys for the scores and over the 64 output dimensions for P V.  Everything after the fp16 q / k / v is fp32.
__global__ void __launch_bounds__(128) clip_causal_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out) {
  __shared__ float kt[CHD][CKP];     // kt[d][j]
  __shared__ float vs[CTOK][CHD];    // vs[j][d]
  __shared__ float qs[4][CHD];
  __shared__ float ps[4][CKP];
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __half* base = qkv + static_cast<size_t>(b) * CTOK * (3 * CH) + h * CHD;
  for (int idx = threadIdx.x; idx < CTOK * (CHD / 8); idx += blockDim.x) {
    const int j = idx / (CHD / 8), v8 = idx % (CHD / 8);
    const uint4 ku = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(j) * (3 * CH) + CH) + v8);
    const uint4 vu = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(j) * (3 * CH) + 2 * CH) + v8);
    const __half* kh = reinterpret_cast<const __half*>(&ku);
    const __half* vh = reinterpret_cast<const __half*>(&vu);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      kt[v8 * 8 + e][j] = __half2float(kh[e]);
      vs[j][v8 * 8 + e] = __half2float(vh[e]);
    }
  }
  __syncthreads();
  for (int i = warp; i < CTOK; i += 4) {
    const __half* qrow = base + static_cast<size_t>(i) * (3 * CH);
    qs[warp][lane] = __half2float(qrow[lane]);
    qs[warp][lane + 32] = __half2float(qrow[lane + 32]);
    __syncwarp();
    float s[3];
    float mx = -INFINITY;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int j = lane + 32 * r;
      float acc = -INFINITY;
      if (j <= i) {  // causal mask: a token attends to itself and the tokens before it
        acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < CHD; ++d) acc = fmaf(qs[warp][d], kt[d][j], acc);
      }
      s[r] = acc;
      mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int j = lane + 32 * r;
      const float e = (j <= i) ? __expf(s[r] - mx) : 0.f;
      if (j < CKP) ps[warp][j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j <= i; ++j) {
      const float pj = ps[warp][j];
      o0 = fmaf(pj, vs[j][lane], o0);
      o1 = fmaf(pj, vs[j][lane + 32], o1);
    }
    const float inv = 1.0f / sum;
    __half* orow = out + (static_cast<size_t>(b) * CTOK + i) * CH + h * CHD;
    orow[lane] = __float2half(o0 * inv);
    orow[lane + 32] = __float2half(o1 * inv);
    __syncwarp();  // qs / ps are rewritten by the next row of this warp
  }
}

// quick_gelu in place: x * sigmoid(1.702 x)   (transformers activations.py QuickGELUActivation)
__global__ void clip_quick_gelu_kernel(uint4* __restrict__ x, size_t nvec) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    uint4 u = x[i];
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = __half22float2(h[e]);
      f.x = f.x / (1.0f + __expf(-1.702f * f.x));
      f.y = f.y / (1.0f + __expf(-1.702f * f.y));
      h[e] = __floats2half2_rn(f.x, f.y);
    }
    x[i] = u;
  }
}

// final LayerNorm with the fp32 output the callers consume: one warp per row of 768 (three 16-byte vectors per lane)
__global__ void __launch_bounds__(256) clip_final_ln_kernel(const __half* __restrict__ x, int rows,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, float* __restrict__ out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * CH);
  float v[3][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const uint4 u = __ldg(src + i * 32 + lane);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h[e]);
      v[i][2 * e] = f.x;
      v[i][2 * e + 1] = f.y;
      sum += f.x + f.y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.0f / CH);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = v[i][e] - mean;
      var = fmaf(d, d, var);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var * (1.0f / CH) + eps);
  float* dst = out + static_cast<size_t>(row) * CH;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c0 = (i * 32 + lane) * 8;
#pragma unroll
    for (int e = 0; e < 8; e += 4) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0 + e));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c0 + e));
      float4 o;
      o.x = (v[i][e] - mean) * rstd * g.x + bb.x;
      o.y = (v[i][e + 1] - mean) * rstd * g.y + bb.y;
      o.z = (v[i][e + 2] - mean) * rstd * g.z + bb.z;
      o.w = (v[i][e + 3] - mean) * rstd * g.w + bb.w;
      *reinterpret_cast<float4*>(dst + c0 + e) = o;
    }
  }
}

struct CLayer {
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  __half *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;  // [2304,768] (q rows pre-scaled), [768,768], [3072,768], [768,3072]
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
};
struct CPlan {
  std::vector<std::function<int(cudaStream_t)>> ops;
  std::vector<std::unique_ptr<GemmPlan>> gemms;
  std::vector<void*> bufs;
  int* ids = nullptr;      // [B*77]
  float* out32 = nullptr;  // [B*77, 768]
  int launches = 0;
};

}  // namespace
}  // namespace pnp

using namespace pnp;

struct pnp_clip {
  int device = 0, num_sms = 148;
  bool finalized = false;
  std::unordered_map<std::string, std::vector<__half>> host;
  std::vector<void*> allocs;
  int vocab = 0, n_layers = 0;
  __half *tok_emb = nullptr, *pos_emb = nullptr;
  std::vector<CLayer> layers;
  float *lnf_g = nullptr, *lnf_b = nullptr;
  std::map<int, std::unique_ptr<CPlan>> plans;  // by prompt count
  std::vector<int> ids_host;
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

const std::vector<__half>* cneed(pnp_clip* e, const std::string& n, int64_t numel, bool* ok) {
  auto it = e->host.find(n);
  if (it == e->host.end() || (numel > 0 && static_cast<int64_t>(it->second.size()) != numel)) {
    set_last_error("clip: parameter missing or wrong size: " + n);
    *ok = false;
    return nullptr;
  }
  return &it->second;
}
float* cup32(pnp_clip* e, const std::vector<__half>& v, float scale = 1.0f) {
  std::vector<float> f(v.size());
  for (size_t i = 0; i < v.size(); ++i) f[i] = __half2float(v[i]) * scale;
  float* d = e->dalloc<float>(f.size());
  if (d && cudaMemcpy(d, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  return d;
}
__half* cup16(pnp_clip* e, const __half* v, size_t n) {
  __half* d = e->dalloc<__half>(n);
  if (d && cudaMemcpy(d, v, n * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  return d;
}

int clip_finalize(pnp_clip* e) {
  bool ok = true;
  const std::string tm = "text_model.";
  const auto* tok = cneed(e, tm + "embeddings.token_embedding.weight", -1, &ok);
  const auto* pos = cneed(e, tm + "embeddings.position_embedding.weight", static_cast<int64_t>(CTOK) * CH, &ok);
  if (!ok) return -2;
  PNP_CHECK(tok->size() % CH == 0 && tok->size() >= CH, "clip: token embedding is not [vocab, 768]");
  e->vocab = static_cast<int>(tok->size() / CH);
  e->tok_emb = cup16(e, tok->data(), tok->size());
  e->pos_emb = cup16(e, pos->data(), pos->size());
  PNP_CHECK(e->tok_emb && e->pos_emb, "clip: upload failed (out of memory?)");
  int nl = 0;
  while (e->host.count(tm + "encoder.layers." + std::to_string(nl) + ".layer_norm1.weight")) ++nl;
  PNP_CHECK(nl >= 1 && nl <= 48, "clip: no encoder layers found (text_model.encoder.layers.<i>.*)");
  e->n_layers = nl;
  for (int l = 0; l < nl; ++l) {
    const std::string p = tm + "encoder.layers." + std::to_string(l) + ".";
    CLayer L;
    const auto* g1 = cneed(e, p + "layer_norm1.weight", CH, &ok);
    const auto* b1 = cneed(e, p + "layer_norm1.bias", CH, &ok);
    const auto* g2 = cneed(e, p + "layer_norm2.weight", CH, &ok);
    const auto* b2 = cneed(e, p + "layer_norm2.bias", CH, &ok);
    const auto* wq = cneed(e, p + "self_attn.q_proj.weight", static_cast<int64_t>(CH) * CH, &ok);
    const auto* wk = cneed(e, p + "self_attn.k_proj.weight", static_cast<int64_t>(CH) * CH, &ok);
    const auto* wv = cneed(e, p + "self_attn.v_proj.weight", static_cast<int64_t>(CH) * CH, &ok);
    const auto* wo = cneed(e, p + "self_attn.out_proj.weight", static_cast<int64_t>(CH) * CH, &ok);
    const auto* bq = cneed(e, p + "self_attn.q_proj.bias", CH, &ok);
    const auto* bk = cneed(e, p + "self_attn.k_proj.bias", CH, &ok);
    const auto* bv = cneed(e, p + "self_attn.v_proj.bias", CH, &ok);
    const auto* bo = cneed(e, p + "self_attn.out_proj.bias", CH, &ok);
    const auto* w1 = cneed(e, p + "mlp.fc1.weight", static_cast<int64_t>(CFF) * CH, &ok);
    const auto* c1 = cneed(e, p + "mlp.fc1.bias", CFF, &ok);
    const auto* w2 = cneed(e, p + "mlp.fc2.weight", static_cast<int64_t>(CH) * CFF, &ok);
    const auto* c2 = cneed(e, p + "mlp.fc2.bias", CH, &ok);
    if (!ok) return -2;
    // fused [q | k | v] projection; the 1/sqrt(64) = 0.125 query scale (modeling_clip.py: `query_states = q_proj(x) *
    // self.scale`) is folded into the q rows and bias: a power of two, so the fp16 weights stay exact
    std::vector<__half> qkv(static_cast<size_t>(3) * CH * CH);
    for (size_t i = 0; i < static_cast<size_t>(CH) * CH; ++i) qkv[i] = __float2half(__half2float((*wq)[i]) * 0.125f);
    memcpy(static_cast<void*>(qkv.data() + static_cast<size_t>(CH) * CH), wk->data(), sizeof(__half) * CH * CH);
    memcpy(static_cast<void*>(qkv.data() + static_cast<size_t>(2) * CH * CH), wv->data(), sizeof(__half) * CH * CH);
    std::vector<float> bqkv(3 * CH);
    for (int i = 0; i < CH; ++i) {
      bqkv[i] = __half2float((*bq)[i]) * 0.125f;
      bqkv[CH + i] = __half2float((*bk)[i]);
      bqkv[2 * CH + i] = __half2float((*bv)[i]);
    }
    L.ln1_g = cup32(e, *g1); L.ln1_b = cup32(e, *b1); L.ln2_g = cup32(e, *g2); L.ln2_b = cup32(e, *b2);
    L.wqkv = cup16(e, qkv.data(), qkv.size());
    L.bqkv = e->dalloc<float>(bqkv.size());
    if (L.bqkv) PNP_CUDA(cudaMemcpy(L.bqkv, bqkv.data(), bqkv.size() * sizeof(float), cudaMemcpyHostToDevice));
    L.wo = cup16(e, wo->data(), wo->size()); L.bo = cup32(e, *bo);
    L.w1 = cup16(e, w1->data(), w1->size()); L.b1 = cup32(e, *c1);
    L.w2 = cup16(e, w2->data(), w2->size()); L.b2 = cup32(e, *c2);
    PNP_CHECK(L.ln1_g && L.ln1_b && L.ln2_g && L.ln2_b && L.wqkv && L.bqkv && L.wo && L.bo && L.w1 && L.b1 && L.w2 && L.b2,
              "clip: upload failed (out of memory?)");
    e->layers.push_back(L);
  }
  const auto* gf = cneed(e, tm + "final_layer_norm.weight", CH, &ok);
  const auto* bf = cneed(e, tm + "final_layer_norm.bias", CH, &ok);
  if (!ok) return -2;
  e->lnf_g = cup32(e, *gf);
  e->lnf_b = cup32(e, *bf);
  PNP_CHECK(e->lnf_g && e->lnf_b, "clip: upload failed (out of memory?)");
  e->host.clear();
  e->finalized = true;
  return 0;
}

int build_clip_plan(pnp_clip* e, int B, CPlan* pl) {
  const int M = B * CTOK;
  int rc = 0;
  auto buf16 = [&](size_t n) -> __half* {
    void* p = nullptr;
    if (cudaMalloc(&p, n * sizeof(__half)) != cudaSuccess) { rc = -1; set_last_error("clip: cudaMalloc failed"); return nullptr; }
    pl->bufs.push_back(p);
    return static_cast<__half*>(p);
  };
  {
    void* p = nullptr;
    PNP_CUDA(cudaMalloc(&p, static_cast<size_t>(M) * sizeof(int)));
    pl->bufs.push_back(p);
    pl->ids = static_cast<int*>(p);
    PNP_CUDA(cudaMalloc(&p, static_cast<size_t>(M) * CH * sizeof(float)));
    pl->bufs.push_back(p);
    pl->out32 = static_cast<float*>(p);
  }
  __half* X = buf16(static_cast<size_t>(M) * CH);
  __half* Y = buf16(static_cast<size_t>(M) * CH);
  __half* NRM = buf16(static_cast<size_t>(M) * CH);
  __half* QKV = buf16(static_cast<size_t>(M) * 3 * CH);
  __half* ATT = buf16(static_cast<size_t>(M) * CH);
  __half* FF = buf16(static_cast<size_t>(M) * CFF);
  if (rc) return rc;
  auto op = [&](int n, std::function<int(cudaStream_t)> f) {
    pl->ops.push_back(std::move(f));
    pl->launches += n;
  };
  auto linear = [&](const __half* in, int K, const __half* wt, const float* bias, int N, __half* out,
                    const __half* residual) {
    if (rc) return;
    GemmEpilogue ep;
    ep.bias = bias;
    ep.out = out;
    ep.ldc = N;
    ep.residual = residual;
    ep.ldr = N;
    ASource s{in, K, K};
    auto g = std::make_unique<GemmPlan>();
    rc = gemm_plan_create(g.get(), &s, 1, 1, true, 1, 1, M, wt, N, K, ep, 0, e->num_sms);
    if (rc) return;
    GemmPlan* gp = g.get();
    pl->gemms.push_back(std::move(g));
    op(1, [gp](cudaStream_t st) { return gemm_launch(*gp, st); });
  };
  {
    const int* ids = pl->ids;
    const __half* tok = e->tok_emb;
    const __half* pos = e->pos_emb;
    op(1, [=](cudaStream_t s) {
      clip_embed_kernel<<<M, 96, 0, s>>>(ids, tok, pos, X);
      PNP_CUDA(cudaGetLastError());
      return 0;
    });
  }
  for (int l = 0; l < e->n_layers; ++l) {
    const CLayer& L = e->layers[l];
    // modeling_clip.py CLIPEncoderLayer.forward: x = x + attn(ln1(x)); x = x + mlp(ln2(x))
    op(1, [=](cudaStream_t s) { return layernorm_launch(X, M, CH, L.ln1_g, L.ln1_b, 1e-5f, NRM, s); });
    linear(NRM, CH, L.wqkv, L.bqkv, 3 * CH, QKV, nullptr);
    op(1, [=](cudaStream_t s) {
      clip_causal_attn_kernel<<<dim3(CHEADS, B), 128, 0, s>>>(QKV, ATT);
      PNP_CUDA(cudaGetLastError());
      return 0;
    });
    linear(ATT, CH, L.wo, L.bo, CH, Y, X);
    op(1, [=](cudaStream_t s) { return layernorm_launch(Y, M, CH, L.ln2_g, L.ln2_b, 1e-5f, NRM, s); });
    linear(NRM, CH, L.w1, L.b1, CFF, FF, nullptr);
    op(1, [=](cudaStream_t s) {
      const size_t nvec = static_cast<size_t>(M) * CFF / 8;
      const int blocks = static_cast<int>(std::min<size_t>((nvec + 255) / 256, 148 * 8));
      clip_quick_gelu_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<uint4*>(FF), nvec);
      PNP_CUDA(cudaGetLastError());
      return 0;
    });
    linear(FF, CFF, L.w2, L.b2, CH, X, Y);
  }
  if (rc) return rc;
  {
    float* out = pl->out32;
    const float* g = e->lnf_g;
    const float* b = e->lnf_b;
    op(1, [=](cudaStream_t s) {
      clip_final_ln_kernel<<<(M * 32 + 255) / 256, 256, 0, s>>>(X, M, g, b, 1e-5f, out);
      PNP_CUDA(cudaGetLastError());
      return 0;
    });
  }
  // split-K workspace shared by the GEMMs of the plan (they run back to back on one stream)
  size_t ws = 0;
  for (auto& g : pl->gemms) ws = std::max(ws, gemm_ws_floats(*g));
  if (ws > 0) {
    float* w = nullptr;
    int* c = nullptr;
    PNP_CUDA(cudaMalloc(reinterpret_cast<void**>(&w), ws * sizeof(float)));
    pl->bufs.push_back(w);
    PNP_CUDA(cudaMalloc(reinterpret_cast<void**>(&c), kGemmMaxCounters * sizeof(int)));
    pl->bufs.push_back(c);
    PNP_CUDA(cudaMemset(c, 0, kGemmMaxCounters * sizeof(int)));
    for (auto& g : pl->gemms) gemm_set_workspace(g.get(), w, c);
  }
  return 0;
}

}  // namespace
}  // namespace pnp

extern "C" {

int pnp_clip_create(int device_ordinal, pnp_clip** out) {
  PNP_CHECK(out != nullptr, "pnp_clip_create: out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e0 = cudaGetDeviceCount(&ndev);
  if (e0 != cudaSuccess || ndev == 0) {
    set_last_error(std::string("pnp_clip_create: no CUDA device available (") + cudaGetErrorString(e0) +
                   "); this library has no CPU fallback");
    return -1;
  }
  PNP_CHECK(device_ordinal >= 0 && device_ordinal < ndev, "pnp_clip_create: bad device ordinal");
  PNP_CUDA(cudaSetDevice(device_ordinal));
  cudaDeviceProp prop;
  PNP_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
  PNP_CHECK(prop.major == 10, "pnp_clip_create: this library is built for sm_100a (B200) only");
  auto* e = new pnp_clip();
  e->device = device_ordinal;
  e->num_sms = prop.multiProcessorCount;
  *out = e;
  return 0;
}

void pnp_clip_destroy(pnp_clip* h) {
  if (h == nullptr) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->plans)
    for (void* p : kv.second->bufs) cudaFree(p);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

int pnp_clip_load_param(pnp_clip* h, const char* name, const uint16_t* data_host, int64_t numel) {
  PNP_CHECK(h && name && data_host && numel > 0, "pnp_clip_load_param: bad argument");
  PNP_CHECK(!h->finalized, "pnp_clip_load_param: parameters already finalized");
  std::vector<__half> v(static_cast<size_t>(numel));
  memcpy(static_cast<void*>(v.data()), data_host, static_cast<size_t>(numel) * sizeof(uint16_t));
  h->host[name] = std::move(v);
  return 0;
}

int pnp_clip_finalize(pnp_clip* h) {
  PNP_CHECK(h != nullptr && !h->finalized, "pnp_clip_finalize: bad handle");
  PNP_CUDA(cudaSetDevice(h->device));
  return clip_finalize(h);
}

int pnp_clip_vocab_size(pnp_clip* h, int* vocab_out, int* layers_out) {
  PNP_CHECK(h && h->finalized && vocab_out && layers_out, "pnp_clip_vocab_size: bad argument");
  *vocab_out = h->vocab;
  *layers_out = h->n_layers;
  return 0;
}

int pnp_clip_encode(pnp_clip* h, const int32_t* input_ids_host, int batch, float* out_dev, void* stream) {
  PNP_CHECK(h && h->finalized && input_ids_host && out_dev, "pnp_clip_encode: bad argument");
  PNP_CHECK(batch >= 1 && batch <= 64, "pnp_clip_encode: 1..64 prompts per call");
  const int M = batch * CTOK;
  for (int i = 0; i < M; ++i)
    if (input_ids_host[i] < 0 || input_ids_host[i] >= h->vocab) {
      set_last_error("pnp_clip_encode: token id " + std::to_string(input_ids_host[i]) + " at position " + std::to_string(i) +
                     " is outside the vocabulary [0, " + std::to_string(h->vocab) + ")");
      return -2;
    }
  PNP_CUDA(cudaSetDevice(h->device));
  auto it = h->plans.find(batch);
  if (it == h->plans.end()) {
    auto pl = std::make_unique<CPlan>();
    int rc = build_clip_plan(h, batch, pl.get());
    if (rc) {
      for (void* p : pl->bufs) cudaFree(p);
      return rc;
    }
    it = h->plans.emplace(batch, std::move(pl)).first;
  }
  CPlan* pl = it->second.get();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // the ids are a few hundred bytes: a pageable copy is staged by the driver before the call returns
  PNP_CUDA(cudaMemcpyAsync(pl->ids, input_ids_host, static_cast<size_t>(M) * sizeof(int), cudaMemcpyHostToDevice, s));
  for (auto& f : pl->ops) {
    int rc = f(s);
    if (rc) return rc;
  }
  h->launches += pl->launches;
  PNP_CUDA(cudaMemcpyAsync(out_dev, pl->out32, static_cast<size_t>(M) * CH * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}

int pnp_clip_kernel_launches(pnp_clip* h, int64_t* out) {
  PNP_CHECK(h && out, "null argument");
  *out = h->launches;
  return 0;
}

}  // extern "C"
