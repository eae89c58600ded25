// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M,N] = A[M,K] . Wt[N,K]^T   fp16 operands, fp32 accumulation in TMEM, fused epilogue.
//
// One persistent CTA per SM, 384 threads, warp-specialised (each role loop runs on a converged warp, one elected lane
// issues):
//   warp 0           TMA producer: per 64-wide K block one 4-D box of the NHWC activation (128 pixels x 64 channels,
//                    shifted by the 3x3 tap, hardware zero fill = the conv padding) and one 2-D box per accumulator of
//                    the weights, all landing 128B-swizzled in a STAGES-deep shared-memory ring (full/empty mbarriers).
//   warp 1           MMA issuer: 4 x NSUB tcgen05.mma (M=128, N=BN, K=16) per stage into NSUB interleaved TMEM
//                    accumulators; tcgen05.commit releases the stage / publishes the accumulators.
//   warp 2           TMEM allocator.
//   warps 4..11      epilogue (two warps per TMEM lane quarter): tcgen05.ld 32 lanes x 32 columns, + bias + time
//                    embedding + residual (or GEGLU, or the fp32 NCHW planes of conv_out), fp16 pack, 256-bit global
//                    stores; overlaps the next tile's MMAs when TMEM holds two accumulator sets.
// Tile shapes: 128 x {64,128,160,256} with one accumulator, 128 x 320 as two accumulators of 160 (NSUB = 2), and an
// opt-in pair mode (PAIR: two CTAs of a cluster, tcgen05.mma.cta_group::2, 256 x BN).  Split-K over the K blocks with a
// deterministic last-arrival reduction.  Optional in-kernel role cycle counters (GemmParams::prof).
//
// This is the only place the library does dense contractions: ResnetBlock2D convs (reference arithmetic:
// models/edict/my_diffusers/models/resnet.py:331-365), 1x1 proj_in/out, attention projections and the GEGLU
// feed-forward (models/edict/my_diffusers/models/attention.py:140-151,186-200,253-260,329-333).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "pnp_internal.h"
#include "pnp_ptx.cuh"

namespace pnp {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;

// PAIR: two CTAs (one thread-block cluster, two SMs) compute one 256 x BN tile with tcgen05.mma.cta_group::2.  Each CTA
// stages its own 128 rows of A and only HALF of the weight tile; the tensor cores read the other half from the peer's
// shared memory.  What bounds these GEMMs is the ~64 B/cycle one SM can pull in from L2, so halving the weight bytes
// per SM raises the arithmetic intensity per ingested byte from 73 to 100 (BN=160) / 85 to 131 FLOP/B (BN=256).
//
// NSUB: the CTA tile is 128 x (NSUB*BN); every K=16 step issues NSUB MMAs of N=BN into NSUB different accumulators.
// Measured (in-kernel cycle counters, profiles/README.md): back-to-back MMAs that accumulate into the SAME TMEM tile
// issue no faster than one per ~112-128 cycles whatever N is, so a lone accumulator runs the tensor pipe at full rate
// only for N=256; two interleaved accumulators of N=160 (a 128x320 tile) hide that latency and read the A tile once
// for twice the columns.
template <int BN, int NSUB, bool PAIR>
struct Cfg {
  static constexpr int BNT = BN * NSUB;                            // columns of the CTA tile
  static constexpr int B_BYTES = (PAIR ? BNT / 2 : BNT) * BK * 2;  // bytes of the weight tile staged by THIS CTA
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = (225 * 1024) / STAGE > 8 ? 8 : (225 * 1024) / STAGE;
  static constexpr int NACC = 2 * BNT <= 512 ? 2 : 1;  // accumulator sets in TMEM (2 = epilogue overlaps the next tile)
  static constexpr int ACC_COLS = NACC * BNT;
  static constexpr int TMEM_COLS = ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512));
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM = STAGES * STAGE + BAR_BYTES + 1024;  // +1024: manual alignment slack
};

// Exact (erf) GELU, branch-free: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the fp16 output rounding),
//   erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2),  t = 1 / (1 + 0.3275911 z),  z = |x| / sqrt(2)
//   gelu(x) = x * (x >= 0 ? 1 - q : q),  q = (poly * exp(-z^2)) / 2
// 14 FP instructions + MUFU.RCP + MUFU.EX2 instead of erff()'s two polynomial branches (the GEGLU epilogue was the
// bottleneck of the FF1 GEMMs: 51 us for 26.8 GFLOP).
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // one MUFU.RCP (2 ulp); __frcp_rn adds a Newton step + fix-up
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  poly = fmaf(t, poly, 0.5f * 1.421413741f);
  poly = fmaf(t, poly, 0.5f * -0.284496736f);
  poly = fmaf(t, poly, 0.5f * 0.254829592f);
  poly *= t;
  const float q = poly * ex2_approx(z * z * -1.4426950408889634f);
  return x * (x >= 0.f ? 1.0f - q : q);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int BN, int NSUB, bool PAIR>
__global__ void __launch_bounds__(384, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<BN, NSUB, PAIR>;
  constexpr int BNT = C::BNT;
  static_assert(!PAIR || NSUB == 1, "pair mode has one accumulator per tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  volatile int* split_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a[0]);
    tma_prefetch_desc(&p.map_b);
    if (p.chunks1 > 0) tma_prefetch_desc(&p.map_a[1]);
    if (p.chunks2 > 0) tma_prefetch_desc(&p.map_a[2]);
    if (PAIR) tma_prefetch_desc(&p.map_b_half);
  }
  const int crank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);  // PAIR: the leader's commit is multicast to this offset in both CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], PAIR ? 16 : 8);  // one arrival per epilogue warp; PAIR: both CTAs' warps arrive on the leader
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_cg2(tmem_slot, C::TMEM_COLS); else tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers and TMEM exist before any copy / commit / MMA reaches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // set-up done: let the successor start its own, then wait for the predecessor's results

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int num_kb = p.num_kb;
  // work items: (tile, K split) for a lone CTA; (two vertically adjacent M tiles, same N tile, K split) for a pair
  const int half_m = p.m_tiles >> 1;
  const int total_work = PAIR ? half_m * p.n_tiles * p.splits : total_tiles * p.splits;
  const int work_begin = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int work_stride = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  auto decode = [&](int work, int& tile, int& split, int& m_blk, int& n_blk) {
    if (PAIR) {
      const int pairs = half_m * p.n_tiles;
      split = work / pairs;
      const int pw = work - split * pairs;
      n_blk = pw / half_m;
      m_blk = 2 * (pw - n_blk * half_m) + crank;
      tile = n_blk * p.m_tiles + m_blk;
    } else {
      tile = work % total_tiles;
      split = work / total_tiles;
      if (p.raster) {
        n_blk = tile % p.n_tiles;
        m_blk = tile / p.n_tiles;
      } else {
        m_blk = tile % p.m_tiles;
        n_blk = tile / p.m_tiles;
      }
    }
  };

  if (warp == 0) {
    {
      // ------------------------------------------------------------ TMA producer
      // The WHOLE warp runs the loop (uniform control flow, operands in uniform registers) and one elected lane issues
      // the copies.  With the loop inside `if (lane == 0)` the compiler treats every operand as divergent and wraps each
      // UTMALDG / UTCHMMA / UTCBAR in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop: ~570 cycles of serial latency per
      // 64-wide K block, more than the tensor pipe needs for any tile narrower than 256 columns.
      uint32_t stage = 0, phase = 0;
      const int kb0 = p.taps0 * p.chunks0;
      const bool prof = p.prof != nullptr;
      long long pr_t0 = prof ? clock64() : 0, pr_wait = 0;
      for (int work = work_begin; work < total_work; work += work_stride) {
        int tile, split, m_blk, n_blk;
        decode(work, tile, split, m_blk, n_blk);
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(num_kb, kb_begin + p.kb_per_split);
        const int p0 = m_blk * BM;
        int x0, y0, b0;
        if (p.linear) {
          x0 = p0; y0 = 0; b0 = 0;
        } else {
          b0 = p0 / p.HW;
          const int rem = p0 - b0 * p.HW;
          y0 = rem / p.W;
          x0 = rem - y0 * p.W;  // 0 unless the image is wider than one tile (W = 256 / 512: the VAE levels)
        }
        // feature injection: the taps of source 0 come from another batch row (the shortcut sources do not)
        const int b0_src0 = (p.a0_row_map != nullptr && !p.linear) ? p.a0_row_map[b0] : b0;
        // K-block cursor kept incrementally (segment, tap offsets, channel chunk): an integer division per K block put
        // a ~150-cycle dependent chain into this loop, which has nothing else to overlap it with
        int seg, cc, dx = 0, dy = 0, tap = 0;
        if (kb_begin < kb0) {
          seg = 0;
          tap = kb_begin / p.chunks0;
          cc = kb_begin - tap * p.chunks0;
          if (p.taps0 == 9) {
            dy = tap / 3 - 1;
            dx = tap - (tap / 3) * 3 - 1;
          }
        } else if (kb_begin - kb0 < p.chunks1) {
          seg = 1;
          cc = kb_begin - kb0;
        } else {
          seg = 2;
          cc = kb_begin - kb0 - p.chunks1;
        }
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          // operands of this K block, computed before the wait
          const CUtensorMap* am = &p.map_a[seg];
          const int c0 = cc * BK, cx = x0 + (seg == 0 ? dx : 0), cy = y0 + (seg == 0 ? dy : 0);
          const int cb = seg == 0 ? b0_src0 : b0;
          ++cc;
          if (seg == 0) {
            if (cc == p.chunks0) {
              cc = 0;
              ++tap;
              if (++dx == 2) { dx = -1; ++dy; }
              if (tap == p.taps0) { seg = 1; dx = dy = 0; }
            }
          } else if (seg == 1 && cc == p.chunks1) {
            seg = 2;
            cc = 0;
          }
          if (prof) {
            const long long t = clock64();
            mbar_wait(&empty[stage], phase ^ 1u, p.dbg, 1);
            pr_wait += clock64() - t;
          } else {
            mbar_wait(&empty[stage], phase ^ 1u, p.dbg, 1);
          }
          uint8_t* sa = smem + stage * C::STAGE;
          uint8_t* sb = sa + A_BYTES;
          if (p.exp & 1) {  // experiment: no copies at all (the MMAs read whatever is in shared memory)
            if ((!PAIR || crank == 0) && lane == 0) mbar_arrive(&full[stage]);
            __syncwarp();
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
            continue;
          }
          if (elect_one()) {
            // PAIR: the boxes of BOTH CTAs complete on the leader's barrier (only the leader's MMA warp waits on it)
            if (!PAIR) mbar_arrive_expect_tx(&full[stage], C::STAGE);
            else if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C::STAGE);
            if (PAIR) {
              tma_load_4d_cg2(sa, am, &full[stage], c0, cx, cy, cb);
              tma_load_2d_cg2(sb, &p.map_b_half, &full[stage], kb * BK, n_blk * BNT + crank * (BNT / 2));
            } else {
              tma_load_4d(sa, am, &full[stage], c0, cx, cy, cb);
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub)  // one box (<= 256 rows) per accumulator
                tma_load_2d(sb + sub * BN * BK * 2, &p.map_b, &full[stage], kb * BK, n_blk * BNT + sub * BN);
            }
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (prof && lane == 0) {
        p.prof[blockIdx.x * 8 + 3] = clock64() - pr_t0;
        p.prof[blockIdx.x * 8 + 4] = pr_wait;
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // ------------------------------------------------------------ MMA issuer (PAIR: the leader CTA drives both SMs)
      // whole warp in the loop, one elected lane issues MMAs and commits (see the producer)
      constexpr uint32_t idesc = umma_idesc_f16(PAIR ? 2 * BM : BM, BN);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      const bool prof = p.prof != nullptr;
      long long mm_t0 = prof ? clock64() : 0, mm_wfull = 0, mm_wtempty = 0;
      for (int work = work_begin; work < total_work; work += work_stride, ++it) {
        int tile, split, m_blk, n_blk;
        decode(work, tile, split, m_blk, n_blk);
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(num_kb, kb_begin + p.kb_per_split);
        const uint32_t as = C::NACC == 2 ? (it & 1) : 0;
        const uint32_t aphase = C::NACC == 2 ? ((it >> 1) & 1) : (it & 1);
        if (prof) {
          const long long t = clock64();
          mbar_wait(&tempty[as], aphase ^ 1u, p.dbg, 2);
          mm_wtempty += clock64() - t;
        } else {
          mbar_wait(&tempty[as], aphase ^ 1u, p.dbg, 2);
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BNT;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (prof) {
            const long long t = clock64();
            mbar_wait(&full[stage], phase, p.dbg, 3);
            mm_wfull += clock64() - t;
          } else {
            mbar_wait(&full[stage], phase, p.dbg, 3);
          }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE);
          const uint64_t adesc = umma_desc_sw128_kmajor(a_addr);
          const uint64_t bdesc = umma_desc_sw128_kmajor(a_addr + A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if ((p.exp & 2) && !(kb == kb_begin && k == 0)) continue;  // experiment: one MMA per tile, copies only
              // +32 bytes per K=16 step inside the 128-byte swizzle atom
              if (PAIR) {
                umma_f16_ss_cg2(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
              } else {
#pragma unroll
                for (int sub = 0; sub < NSUB; ++sub)  // consecutive MMAs go to different accumulators
                  umma_f16_ss(d_tmem + sub * BN, adesc + 2u * k, bdesc + static_cast<uint64_t>(sub * ((BN * BK * 2) >> 4)) + 2u * k,
                              idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
              }
            }
            if (PAIR) umma_commit_mc_cg2(&empty[stage], 0x3); else umma_commit(&empty[stage]);
            if (kb == kb_end - 1) {
              if (PAIR) umma_commit_mc_cg2(&tfull[as], 0x3); else umma_commit(&tfull[as]);
            }
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (prof && lane == 0) {
        p.prof[blockIdx.x * 8 + 0] = clock64() - mm_t0;
        p.prof[blockIdx.x * 8 + 1] = mm_wfull;
        p.prof[blockIdx.x * 8 + 2] = mm_wtempty;
      }
    }
  } else if (warp >= 4) {
    // -------------------------------------------------------------- epilogue (8 warps: 2 per TMEM lane quarter)
    const int q = warp & 3;            // TMEM lane quarter this warp may read (warp id % 4)
    const int half = (warp - 4) >> 2;  // which half of the 32-column chunks this warp owns
    const int row = q * 32 + lane;
    const float* temb = nullptr;
    if (p.temb_table != nullptr) temb = p.temb_table + static_cast<size_t>(*p.t_index) * p.temb_stride;
    int it = 0;
    const bool prof = p.prof != nullptr && threadIdx.x == 128;
    long long ep_t0 = prof ? clock64() : 0, ep_wait = 0;
    for (int work = work_begin; work < total_work; work += work_stride, ++it) {
      int tile, split, m_blk, n_blk;
      decode(work, tile, split, m_blk, n_blk);
      const uint32_t as = C::NACC == 2 ? (it & 1) : 0;
      const uint32_t aphase = C::NACC == 2 ? ((it >> 1) & 1) : (it & 1);
      const int m = m_blk * BM + row;
      const bool valid = m < p.M;
      // the residual does not depend on the accumulator: fetch the first chunk before waiting for the MMAs
      // ... three chunks deep: a K = 320 projection with a residual is HBM-bound and its epilogue thread has one 64-byte
      // row segment per chunk to fetch; with one chunk of look-ahead an SM kept 16 KB of residual in flight (3.4 TB/s over
      // the chip for such a launch), with three it is 48 KB
      constexpr int RD = 3;
      uint4 rb[RD][4];
      const bool has_res = p.residual != nullptr && valid && !p.geglu;
      const __half* res_row = has_res ? p.residual + static_cast<size_t>(m) * p.ldr + n_blk * BNT : nullptr;
#pragma unroll
      for (int d = 0; d < RD; ++d) {
        const int c = half + 2 * d;
        if (has_res && c < BNT / 32) {
          ldg256_nc(res_row + c * 32, rb[d][0], rb[d][1]);
          ldg256_nc(res_row + c * 32 + 16, rb[d][2], rb[d][3]);
        }
      }
      if (prof) {
        const long long t = clock64();
        mbar_wait(&tfull[as], aphase, p.dbg, 4);
        ep_wait += clock64() - t;
      } else {
        mbar_wait(&tfull[as], aphase, p.dbg, 4);
      }
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BNT;
      bool from_ws = false;
      if (p.splits > 1) {
        // dump the raw fp32 partial of this K range, release the accumulator, then elect the last arrival
        float* wrow = p.ws + (static_cast<size_t>(split) * p.M + (valid ? m : 0)) * p.N + n_blk * BNT;
#pragma unroll 1
        for (int c = half; c < BNT / 32; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              stg256_cg(wrow + c * 32 + j,
                        make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                    __uint_as_float(r[j + 3])),
                        make_float4(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]), __uint_as_float(r[j + 6]),
                                    __uint_as_float(r[j + 7])));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(&tempty[as], 0); else mbar_arrive(&tempty[as]);
        }
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 128) {
          const int old = atomicAdd(&p.counters[tile], 1);
          *split_flag = (old == p.splits - 1) ? 1 : 0;
          if (old == p.splits - 1) p.counters[tile] = 0;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const bool last = *split_flag != 0;
        asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone has read the flag before it can be rewritten
        if (!last) continue;
        __threadfence();
        from_ws = true;
      }
      if (!p.geglu) {
#pragma unroll
        for (int c0 = 0; c0 < BNT / 32; c0 += 2) {
          const int c = c0 + half;
          if (c < BNT / 32) {
            const int n0 = n_blk * BNT + c * 32;
            float v[32];
            if (!from_ws) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(taddr + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
              if (valid) {
                // two K slices per round trip (16 independent 16-byte loads in flight); summed in slice order either way
                const size_t slice = static_cast<size_t>(p.M) * p.N;
                const float* wr = p.ws + static_cast<size_t>(m) * p.N + n0;
                int sp = 0;
                for (; sp + 1 < p.splits; sp += 2, wr += 2 * slice) {
                  float4 ta[8], tb[8];
#pragma unroll
                  for (int j = 0; j < 8; j += 2) {
                    ldg256_cg(wr + 4 * j, ta[j], ta[j + 1]);
                    ldg256_cg(wr + slice + 4 * j, tb[j], tb[j + 1]);
                  }
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    v[4 * j] = (v[4 * j] + ta[j].x) + tb[j].x;
                    v[4 * j + 1] = (v[4 * j + 1] + ta[j].y) + tb[j].y;
                    v[4 * j + 2] = (v[4 * j + 2] + ta[j].z) + tb[j].z;
                    v[4 * j + 3] = (v[4 * j + 3] + ta[j].w) + tb[j].w;
                  }
                }
                if (sp < p.splits) {
#pragma unroll
                  for (int j = 0; j < 32; j += 8) {
                    float4 t4, t5;
                    ldg256_cg(wr + j, t4, t5);
                    v[j] += t4.x; v[j + 1] += t4.y; v[j + 2] += t4.z; v[j + 3] += t4.w;
                    v[j + 4] += t5.x; v[j + 5] += t5.y; v[j + 6] += t5.z; v[j + 7] += t5.w;
                  }
                }
              }
            }
            if (p.out_scale != 1.0f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
            }
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
              }
            }
            if (temb != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(temb + n0 + j));
                v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
              }
            }
            if (p.out32 != nullptr) {
              // conv_out: the first four columns go out as fp32 NCHW, one image plane per column (coalesced over pixels)
              if (valid && c == 0) {
                const int bimg = m / p.HW, pix = m - bimg * p.HW;
                float* o32 = p.out32 + static_cast<size_t>(bimg) * p.out32_ch * p.HW + pix;
#pragma unroll
                for (int co = 0; co < 8; ++co)
                  if (co < p.out32_ch) o32[static_cast<size_t>(co) * p.HW] = v[co];
              }
            } else if (valid) {
              if (has_res) {
                const int slot = (c0 / 2) % RD;  // static after unrolling: the ring stays in registers
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const __half2* h = reinterpret_cast<const __half2*>(&rb[slot][j]);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(h[e]);
                    v[j * 8 + e * 2] += f.x;
                    v[j * 8 + e * 2 + 1] += f.y;
                  }
                }
                if (c + 2 * RD < BNT / 32) {  // refill the slot with the chunk three iterations ahead
                  ldg256_nc(res_row + (c + 2 * RD) * 32, rb[slot][0], rb[slot][1]);
                  ldg256_nc(res_row + (c + 2 * RD) * 32 + 16, rb[slot][2], rb[slot][3]);
                }
              }
              __half* op = p.out + static_cast<size_t>(m) * p.ldc + n0;
              uint4 u[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                u[j].x = pack_half2(v[j * 8 + 0], v[j * 8 + 1]);
                u[j].y = pack_half2(v[j * 8 + 2], v[j * 8 + 3]);
                u[j].z = pack_half2(v[j * 8 + 4], v[j * 8 + 5]);
                u[j].w = pack_half2(v[j * 8 + 6], v[j * 8 + 7]);
              }
              stg256(op, u[0], u[1]);
              stg256(op + 16, u[2], u[3]);
            }
          }
        }
      } else {
        // GEGLU: tile columns [0,BNT/2) hold the value projection, [BNT/2,BNT) the gate projection of the same
        // output columns (weights are packed that way by the engine).  attention.py:329-333 (erf GELU).
#pragma unroll 1
        for (int c = half; c < BNT / 64; c += 2) {
          uint32_t rv[32], rg[32];
          tmem_ld_32x32b_x32(taddr + c * 32, rv);
          tmem_ld_32x32b_x32(taddr + BNT / 2 + c * 32, rg);
          tmem_ld_wait();
          const int nv = n_blk * BNT + c * 32;
          const int ng = nv + BNT / 2;
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + nv + j));
            const float4 bg = __ldg(reinterpret_cast<const float4*>(p.bias + ng + j));
            o[j] = (__uint_as_float(rv[j]) + bv.x) * gelu_erf(__uint_as_float(rg[j]) + bg.x);
            o[j + 1] = (__uint_as_float(rv[j + 1]) + bv.y) * gelu_erf(__uint_as_float(rg[j + 1]) + bg.y);
            o[j + 2] = (__uint_as_float(rv[j + 2]) + bv.z) * gelu_erf(__uint_as_float(rg[j + 2]) + bg.z);
            o[j + 3] = (__uint_as_float(rv[j + 3]) + bv.w) * gelu_erf(__uint_as_float(rg[j + 3]) + bg.w);
          }
          if (valid) {
            __half* op = p.out + static_cast<size_t>(m) * p.ldc + n_blk * (BNT / 2) + c * 32;
            uint4 u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              u[j].x = pack_half2(o[j * 8 + 0], o[j * 8 + 1]);
              u[j].y = pack_half2(o[j * 8 + 2], o[j * 8 + 3]);
              u[j].z = pack_half2(o[j * 8 + 4], o[j * 8 + 5]);
              u[j].w = pack_half2(o[j * 8 + 6], o[j * 8 + 7]);
            }
            stg256(op, u[0], u[1]);
            stg256(op + 16, u[2], u[3]);
          }
        }
      }
      if (!from_ws) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(&tempty[as], 0); else mbar_arrive(&tempty[as]);
        }
      }
    }
    if (prof) {
      p.prof[blockIdx.x * 8 + 5] = clock64() - ep_t0;
      p.prof[blockIdx.x * 8 + 6] = ep_wait;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // nobody leaves while the peer may still commit to / arrive on / read from this CTA
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_cg2(tmem_base, C::TMEM_COLS); else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved at run time so that the library has no link-time dependency on libcuda (absent on the build box)
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
  });
  return fn;
}

}  // namespace

int encode_tensor_map_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  PNP_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
             static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    set_last_error(buf);
    return -3;
  }
  return 0;
}

namespace {

template <int BN, int NSUB>
int launch_t(const GemmPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    PNP_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, NSUB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg<BN, NSUB, false>::SMEM));
    if (NSUB == 1)
      PNP_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg<BN, 1, true>::SMEM));
    attr_set = true;
  }
  if (plan.cluster == 2 && NSUB == 1)
    PNP_CUDA(launch_kc(gemm_tcgen05_kernel<BN, 1, true>, dim3(plan.grid), dim3(384), Cfg<BN, 1, true>::SMEM, stream, 2, plan.p));
  else
    PNP_CUDA(launch_k(gemm_tcgen05_kernel<BN, NSUB, false>, dim3(plan.grid), dim3(384), Cfg<BN, NSUB, false>::SMEM, stream,
                      plan.p));
  return 0;
}

}  // namespace

// Cost model (cycles) behind the tile-shape / split-K choice, calibrated with the in-kernel cycle counters
// (PNP_GEMM_PROF, profiles/README.md).  Per CTA and 64-wide K block the slowest of
//   tensor pipe   : 4 K-steps x nsub MMAs; an MMA of N columns takes N/2 cycles, but back-to-back MMAs into the SAME
//                   accumulator issue no faster than ~120 cycles apart, so a lone accumulator costs max(N/2, 120) per step
//   TMA producer  : ~220 cycles of serial issue latency per K block (wait, expect_tx, 1 + nsub copies)
//   operand bytes : (16 KB + tile columns x 128 B) at the ~64 B/cycle one SM pulls from L2
// plus the chip-wide L2 limit, the epilogue (~700 cycles per 32-column chunk pair) and, for split-K, the round trip of
// the fp32 partials.
static long gemm_cost(int M, int N, int num_kb, int bn, int nsub, int splits, int num_sms, bool pair = false) {
  const long bnt = static_cast<long>(bn) * nsub;
  const long m_tiles = (M + BM - 1) / BM;
  const long tiles = m_tiles * (N / bnt);
  const long kb_per = (num_kb + splits - 1) / splits;
  const long ctas = tiles * splits;
  const long waves = (ctas + num_sms - 1) / num_sms;
  // measured MMA-warp cycles per K block: ~480 for one accumulator whatever its width (dependent MMAs ~120 cycles apart),
  // 537 for N=256, 800 for two accumulators of 160 (100 per MMA: shared-memory operand bandwidth)
  const long mma = nsub > 1 ? 4 * nsub * std::max<long>(bn * 5 / 8, 60) : std::max<long>(2 * bn + 25, 480);
  const long stage_l2 = 16384L + (pair ? bnt * 64L : bnt * 128L);
  const long per_kb = std::max<long>(std::max<long>(mma, 300 + 40 * (nsub - 1)), stage_l2 / 64);
  const long tensor = waves * kb_per * per_kb;
  const long l2 = tiles * num_kb * stage_l2 / 5000;
  // epilogue: ~1800 cycles per 64 columns (two warps per lane quarter, 32 columns each); hidden behind the next tile's
  // MMAs when TMEM holds two accumulator sets; split-K adds the fp32 round trip, paid by the last-arriving CTA
  const long chunks = (bnt + 63) / 64;
  long epi = chunks * 1800;
  if (2 * bnt > 512 && waves > 1) epi += (waves - 1) * chunks * 1800;
  if (splits > 1) epi += waves * chunks * 1500 + chunks * 2500L * splits + 6000;
  return std::max(tensor, l2) + epi + 4000;
}

static bool g_cluster_ok = false;  // pair mode (cta_group::2) is opt-in: PNP_GEMM_CLUSTER=1
bool cluster_allowed() { return g_cluster_ok; }
void set_cluster_allowed(bool on) { g_cluster_ok = on; }

long gemm_model_cost(int M, int N, int num_kb, bool geglu, int bnt, int splits, int num_sms) {
  int bn = 0, sp = 0;
  gemm_choose(M, N, num_kb, geglu, num_sms, bnt, splits, &bn, &sp);
  if (bn != bnt || sp != splits) return -1;
  const int nsub = bnt == 320 ? 2 : 1;
  return gemm_cost(M, N, num_kb, bnt / nsub, nsub, splits, num_sms, false);
}

int gemm_choose_bn(int M, int N, bool geglu, int num_sms) {
  int bn = 0, sp = 0;
  gemm_choose(M, N, 64, geglu, num_sms, 0, 1, &bn, &sp);
  return bn;
}

// bn_out encodes the tile: 64/128/160/256 = one accumulator of that width, 320 = two accumulators of 160
void gemm_choose(int M, int N, int num_kb, bool geglu, int num_sms, int bn_force, int split_force, int* bn_out,
                 int* splits_out) {
  const int cands[5][2] = {{160, 2}, {256, 1}, {160, 1}, {128, 1}, {64, 1}};
  long best = -1;
  *bn_out = 0;
  *splits_out = 1;
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i][0], nsub = cands[i][1], bnt = bn * nsub;
    if (bn_force > 0 && bnt != bn_force) continue;
    if (N % bnt != 0) continue;
    if (geglu && bnt != 256 && bnt != 128) continue;
    const long tiles = static_cast<long>((M + BM - 1) / BM) * (N / bnt);
    for (int sp = 1; sp <= 16; ++sp) {
      if (split_force > 0 && sp != split_force) continue;
      if (sp > 1 && (geglu || tiles > kGemmMaxCounters || num_kb / sp < 4)) continue;
      // a split must not leave an empty K range
      const int kb_per = (num_kb + sp - 1) / sp;
      if ((sp - 1) * kb_per >= num_kb) continue;
      const bool pair = cluster_allowed() && nsub == 1 && ((M + BM - 1) / BM) % 2 == 0;
      const long c = gemm_cost(M, N, num_kb, bn, nsub, sp, num_sms, pair);
      if (best < 0 || c < best) {
        best = c;
        *bn_out = bnt;
        *splits_out = sp;
      }
    }
  }
}

int gemm_plan_create(GemmPlan* plan, const ASource* srcs, int nsrc, int taps0, bool linear, int B, int H, int W,
                     const __half* Wt, int N, int Ktot, const GemmEpilogue& ep, int bn_force, int num_sms,
                     int split_force) {
  if (const char* ev = getenv("PNP_GEMM_CLUSTER")) set_cluster_allowed(atoi(ev) != 0);
  PNP_CHECK(nsrc >= 1 && nsrc <= 3, "gemm: 1..3 A sources");
  PNP_CHECK(taps0 == 1 || taps0 == 9, "gemm: taps must be 1 or 9");
  GemmParams& p = plan->p;
  memset(&p, 0, sizeof p);
  const int M = B * H * W;
  int ksum = 0;
  for (int i = 0; i < nsrc; ++i) {
    PNP_CHECK(srcs[i].C % BK == 0, "gemm: every A source needs a multiple of 64 channels");
    PNP_CHECK(srcs[i].ld % 8 == 0 && (reinterpret_cast<uintptr_t>(srcs[i].ptr) & 15) == 0, "gemm: A alignment");
    ksum += srcs[i].C * (i == 0 ? taps0 : 1);
  }
  PNP_CHECK(ksum == Ktot, "gemm: K of the sources does not match the packed weight");
  const bool geglu = ep.geglu;
  int bn = 0, splits = 1;
  gemm_choose(M, N, Ktot / BK, geglu, num_sms, bn_force, split_force, &bn, &splits);
  if (bn == 0 && split_force > 1) gemm_choose(M, N, Ktot / BK, geglu, num_sms, bn_force, 0, &bn, &splits);
  PNP_CHECK(bn != 0, "gemm: no valid tile shape for this N");
  PNP_CHECK(bn == 320 || bn == 256 || bn == 160 || bn == 128 || bn == 64, "gemm: unsupported BN");
  PNP_CHECK(N % bn == 0, "gemm: N must be a multiple of the column tile");
  const int bnt = bn;  // columns of the CTA tile
  const int nsub = bn == 320 ? 2 : 1;
  bn = bnt / nsub;     // N of one MMA = rows of one weight box
  PNP_CHECK(!geglu || bn == 256 || bn == 128, "gemm: GEGLU epilogue needs BN 128/256");
  // the epilogue moves 32 bytes per lane and instruction (256-bit global accesses)
  PNP_CHECK(ep.out_f32_nchw4 != nullptr ||
                (ep.out != nullptr && ep.ldc % 16 == 0 && (reinterpret_cast<uintptr_t>(ep.out) & 31) == 0),
            "gemm: output alignment (32 bytes, ldc % 16 == 0)");
  PNP_CHECK(ep.out_f32_nchw4 == nullptr || (!linear && !geglu && ep.residual == nullptr),
            "gemm: the fp32 NCHW output is for a plain convolution");
  PNP_CHECK(ep.residual == nullptr || (ep.ldr % 16 == 0 && (reinterpret_cast<uintptr_t>(ep.residual) & 31) == 0),
            "gemm: residual alignment (32 bytes, ldr % 16 == 0)");
  PNP_CHECK(!geglu || ep.bias != nullptr, "gemm: GEGLU needs a bias");

  // A maps.  Box = 128 consecutive pixels x 64 channels.
  uint32_t box[4];
  if (linear) {
    PNP_CHECK(B == 1 && H == 1, "gemm: linear mode takes M as W");
    box[0] = BK; box[1] = BM; box[2] = 1; box[3] = 1;
  } else {
    PNP_CHECK((W <= 128 && 128 % W == 0) || W % 128 == 0, "gemm: conv mode needs W | 128 or 128 | W");
    const int rows = 128 / W;  // image rows per tile (0: the image is wider than a tile)
    if (rows == 0) {
      box[0] = BK; box[1] = 128; box[2] = 1; box[3] = 1;
    } else if (rows <= H) {
      PNP_CHECK(H % rows == 0, "gemm: conv tile rows must divide H");
      box[0] = BK; box[1] = W; box[2] = rows; box[3] = 1;
    } else {
      PNP_CHECK(rows % H == 0, "gemm: conv tile must cover whole images");
      box[0] = BK; box[1] = W; box[2] = H; box[3] = rows / H;
    }
  }
  for (int i = 0; i < nsrc; ++i) {
    uint64_t dims[4], strides[3];
    if (linear) {
      dims[0] = srcs[i].C; dims[1] = M; dims[2] = 1; dims[3] = 1;
      strides[0] = static_cast<uint64_t>(srcs[i].ld) * 2;
      strides[1] = strides[0] * M;
      strides[2] = strides[1];
    } else {
      dims[0] = srcs[i].C; dims[1] = W; dims[2] = H; dims[3] = B;
      strides[0] = static_cast<uint64_t>(srcs[i].ld) * 2;
      strides[1] = strides[0] * W;
      strides[2] = strides[1] * H;
    }
    int rc = encode_tensor_map_f16(&p.map_a[i], srcs[i].ptr, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(Ktot) * 2};
    uint32_t bbox[2] = {BK, static_cast<uint32_t>(bn)};
    PNP_CHECK((reinterpret_cast<uintptr_t>(Wt) & 15) == 0 && Ktot % 8 == 0, "gemm: weight alignment");
    int rc = encode_tensor_map_f16(&p.map_b, Wt, 2, dims, strides, bbox);
    if (rc) return rc;
  }
  p.taps0 = taps0;
  p.chunks0 = srcs[0].C / BK;
  p.chunks1 = nsrc > 1 ? srcs[1].C / BK : 0;
  p.chunks2 = nsrc > 2 ? srcs[2].C / BK : 0;
  p.num_kb = Ktot / BK;
  p.linear = linear ? 1 : 0;
  p.W = W;
  p.HW = H * W;
  p.M = M;
  p.N = N;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = N / bnt;
  p.bias = ep.bias;
  p.temb_table = ep.temb_table;
  p.t_index = ep.t_index;
  p.temb_stride = ep.temb_stride;
  p.residual = ep.residual;
  p.ldr = ep.ldr;
  p.out = ep.out;
  p.a0_row_map = ep.a0_row_map;
  PNP_CHECK(ep.a0_row_map == nullptr || (!linear && 128 / W >= 1 && 128 / W <= H),
            "gemm: a source row map needs conv tiles that stay inside one image");
  p.out32 = ep.out_f32_nchw4;
  p.out32_ch = ep.out32_channels;
  p.out_scale = ep.out_scale;
  PNP_CHECK(ep.out32_channels >= 1 && ep.out32_channels <= 8, "gemm: 1..8 fp32 NCHW output planes");
  p.ldc = ep.ldc;
  p.geglu = geglu ? 1 : 0;
  p.dbg = debug_words_device();
  // split-K when the output has too few tiles to occupy the chip (8x8 / 16x16 levels: weight streaming)
  p.splits = splits;
  p.kb_per_split = (p.num_kb + splits - 1) / splits;
  const int tiles = p.m_tiles * p.n_tiles;
  {
    // Tile order.  With M tiles fastest the 148 CTAs in flight share one weight tile and walk down the activations, once
    // per N tile: fine while the activations stay in L2 between passes, but an 84 MB activation (B = 32 at 64x64x320)
    // was streamed from HBM n_tiles times (ncu, GEGLU projection: 754 MB read for 84 MB of A, 54 % DRAM, 49 % tensor).
    // N tiles fastest makes the CTAs in flight share activation tiles instead; the weights (<= 30 MB) live in L2.
    static const int force = [] { const char* e = getenv("PNP_GEMM_RASTER"); return e ? atoi(e) : -1; }();
    size_t a_cols = 0;
    for (int i = 0; i < nsrc; ++i) a_cols += static_cast<size_t>(srcs[i].C);
    const size_t a_bytes = static_cast<size_t>(M) * a_cols * 2;
    static const size_t min_mb = [] { const char* e = getenv("PNP_GEMM_RASTER_MB"); return e ? atoi(e) : 40; }();
    p.raster = force >= 0 ? force : ((p.n_tiles > 1 && a_bytes > (min_mb << 20)) ? 1 : 0);
  }
  plan->bn = bn;
  plan->nsub = nsub;
  {
    size_t a_cols = 0;
    for (int i = 0; i < nsrc; ++i) a_cols += static_cast<size_t>(srcs[i].C);
    const size_t out_bytes = ep.out_f32_nchw4 ? static_cast<size_t>(M) * ep.out32_channels * 4
                                              : static_cast<size_t>(M) * (geglu ? N / 2 : N) * 2;
    plan->algo_bytes = static_cast<size_t>(M) * a_cols * 2 + static_cast<size_t>(N) * Ktot * 2 + out_bytes +
                       (ep.residual ? static_cast<size_t>(M) * N * 2 : 0);
  }
  plan->grid = std::min(tiles * p.splits, num_sms);
  // pair mode (cta_group::2: two vertically adjacent M tiles, each SM stages half of the weight tile) whenever it applies
  plan->cluster = 1;
  if (cluster_allowed() && nsub == 1 && p.m_tiles % 2 == 0) {
    plan->cluster = 2;
    const int pairs = tiles / 2 * p.splits;
    plan->grid = 2 * std::min(pairs, num_sms / 2);
    uint64_t dims[2] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(Ktot) * 2};
    uint32_t hbox[2] = {BK, static_cast<uint32_t>(bn / 2)};
    int rc = encode_tensor_map_f16(&p.map_b_half, Wt, 2, dims, strides, hbox);
    if (rc) return rc;
  }
  return 0;
}

size_t gemm_ws_floats(const GemmPlan& plan) {
  return plan.p.splits > 1 ? static_cast<size_t>(plan.p.splits) * plan.p.M * plan.p.N : 0;
}
void gemm_set_workspace(GemmPlan* plan, float* ws, int* counters) {
  plan->p.ws = ws;
  plan->p.counters = counters;
}

int gemm_launch(const GemmPlan& plan, cudaStream_t stream) {
  if (plan.p.splits > 1 && (plan.p.ws == nullptr || plan.p.counters == nullptr)) {
    set_last_error("gemm_launch: split-K plan without workspace");
    return -2;
  }
  switch (plan.bn * plan.nsub) {
    case 320: return launch_t<160, 2>(plan, stream);
    case 256: return launch_t<256, 1>(plan, stream);
    case 160: return launch_t<160, 1>(plan, stream);
    case 128: return launch_t<128, 1>(plan, stream);
    case 64: return launch_t<64, 1>(plan, stream);
  }
  set_last_error("gemm_launch: bad plan");
  return -2;
}

}  // namespace pnp
