// Micro-benchmark of the tcgen05.mma instruction cost for the shapes the attention kernels use or could use (test entry
// point pnp_test_mma_probe, tools/mma_probe.py).  Question it answers: the d = 40 attention issues M=128 MMAs with N = 48
// (P V) and N = 128 (Q K^T) that cost ~100-110 cycles each whatever N is (profiles/README.md) - is that floor paid per
// INSTRUCTION (then a cta_group::2 instruction covering 256 query rows on two SMs halves it per SM) or per 128 rows of the A
// operand (then it does not), and what do M = 64 instructions (the transposed product O^T = V^T P^T) cost?
//
// One CTA (or one cluster of two for cta_group::2) issues `n` back-to-back MMAs of one shape on whatever bits are in its
// shared / tensor memory, round-robin over `nacc` accumulators, in elected issue blocks of `group` instructions with or
// without a tcgen05.commit after each block, and reports the cycles until the final commit arrives.
#include <cstring>
#include <string>

#include "../../include/pnpinv.h"
#include "pnp_internal.h"
#include "pnp_ptx.cuh"

namespace pnp {
namespace {

struct ProbeParams {
  int cg2, M, N, ts, n, nacc;
  int group, commit;  // MMAs per elected issue block (1..8); 1 = a tcgen05.commit after every block (onto a barrier nobody waits on)
  long long* out;  // [4]: issue cycles, total cycles, n, clock rate placeholder
  volatile unsigned int* dbg;
};

constexpr int PROBE_OPERAND_BYTES = 64 * 1024;
constexpr int PROBE_SMEM = PROBE_OPERAND_BYTES + 256 + 1024;
constexpr int PROBE_COL_A = 448;

template <bool CG2>
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + PROBE_OPERAND_BYTES);
  uint64_t* sink = done + 1;  // target of the intermediate commits: more pending arrivals than commits, never waited on
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = CG2 ? cluster_ctarank() : 0u;
  // operands: fp16 1.0 everywhere (any finite pattern would do)
  for (int i = threadIdx.x; i < PROBE_OPERAND_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    mbar_init(done, 1);
    mbar_init(sink, 0xFFFFFu);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG2) tmem_alloc_cg2(tmem_slot, 512); else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  if (CG2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  {
    // A operand in tensor memory: 8 columns of packed fp16 pairs per K = 16 step, all lanes
    uint32_t ones[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ones[i] = 0x3C003C00u;
    tmem_st_32x32b_x8(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + PROBE_COL_A, ones);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (CG2) cluster_sync_all();
  tc_fence_after();
  if (warp == 0) {
    if (crank == 0) {
      const uint32_t idesc = (1u << 4) | ((static_cast<uint32_t>(p.N) >> 3) << 17) | ((static_cast<uint32_t>(p.M) >> 4) << 24);
      const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem));
      const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem + 32 * 1024));
      const uint32_t stride = static_cast<uint32_t>((p.N + 31) & ~31);
      const long long t0 = clock64();
      for (int it = 0; it < p.n; it += p.group) {
        if (elect_one()) {
#pragma unroll 1
          for (int k = 0; k < p.group; ++k) {
            const int i = it + k;
            const uint32_t d = tmem_base + static_cast<uint32_t>(i % p.nacc) * stride;
            const uint32_t acc = i >= p.nacc ? 1u : 0u;
            // successive K = 16 steps inside one 128-byte swizzle atom, like the product kernels (k & 3)
            const uint64_t bd = bdesc + 2u * (k & 3);
            if (CG2) {
              if (p.ts) umma_f16_ts_cg2(d, tmem_base + PROBE_COL_A, bd, idesc, acc);
              else umma_f16_ss_cg2(d, adesc + 2u * (k & 3), bd, idesc, acc);
            } else {
              if (p.ts) umma_f16_ts(d, tmem_base + PROBE_COL_A, bd, idesc, acc);
              else umma_f16_ss(d, adesc + 2u * (k & 3), bd, idesc, acc);
            }
          }
          if (p.commit) {  // what the product kernels do after every K block / S tile / P V product
            if (CG2) umma_commit_mc_cg2(sink, 0x3); else umma_commit(sink);
          }
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (elect_one()) {
        if (CG2) umma_commit_mc_cg2(done, 0x3); else umma_commit(done);
      }
      __syncwarp();
      mbar_wait(done, 0, p.dbg, 77);
      const long long t2 = clock64();
      if (lane == 0) {
        p.out[0] = t1 - t0;
        p.out[1] = t2 - t0;
        p.out[2] = p.n;
      }
    } else {
      mbar_wait(done, 0, p.dbg, 78);  // the leader's commit is multicast to the peer as well
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (CG2) tmem_dealloc_cg2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace
}  // namespace pnp

using namespace pnp;

extern "C" int pnp_test_mma_probe(int cta_group, int M, int N, int a_from_tmem, int n, int nacc, int group, int commit,
                                  int64_t* cycles_out_host) {
  PNP_CHECK(cycles_out_host != nullptr, "pnp_test_mma_probe: null output");
  PNP_CHECK(cta_group == 1 || cta_group == 2, "pnp_test_mma_probe: cta_group 1 or 2");
  PNP_CHECK(group >= 1 && group <= 8, "pnp_test_mma_probe: 1..8 MMAs per issue block");
  PNP_CHECK(n >= 8 && n % group == 0 && n <= (1 << 18), "pnp_test_mma_probe: n must be a multiple of the block size, <= 2^18");
  PNP_CHECK(M == 64 || M == 128 || M == 256, "pnp_test_mma_probe: M");
  PNP_CHECK((cta_group == 2) == (M == 256), "pnp_test_mma_probe: M = 256 needs cta_group 2 (and only that is probed)");
  PNP_CHECK(N >= 16 && N <= 256 && N % 16 == 0, "pnp_test_mma_probe: N");
  PNP_CHECK(!(cta_group == 2 && a_from_tmem) || N % 32 == 0, "pnp_test_mma_probe: cta_group::2 with A in TMEM needs N % 32 == 0");
  PNP_CHECK(!(M == 64 && a_from_tmem), "pnp_test_mma_probe: M = 64 is probed with A in shared memory");
  PNP_CHECK(nacc >= 1 && nacc * ((N + 31) & ~31) <= PROBE_COL_A, "pnp_test_mma_probe: accumulators do not fit tensor memory");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_last_error("pnp_test_mma_probe: no CUDA device available; this library has no CPU fallback");
    return -1;
  }
  long long* out = nullptr;
  PNP_CUDA(cudaMalloc(reinterpret_cast<void**>(&out), 4 * sizeof(long long)));
  PNP_CUDA(cudaMemset(out, 0, 4 * sizeof(long long)));
  ProbeParams p;
  p.cg2 = cta_group == 2;
  p.M = M;
  p.N = N;
  p.ts = a_from_tmem;
  p.n = n;
  p.nacc = nacc;
  p.group = group;
  p.commit = commit ? 1 : 0;
  p.out = out;
  p.dbg = debug_words_device();
  cudaError_t e;
  if (cta_group == 2) {
    PNP_CUDA(cudaFuncSetAttribute(mma_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PROBE_SMEM));
    e = launch_kc(mma_probe_kernel<true>, dim3(2), dim3(128), PROBE_SMEM, nullptr, 2, p);
  } else {
    PNP_CUDA(cudaFuncSetAttribute(mma_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PROBE_SMEM));
    e = launch_k(mma_probe_kernel<false>, dim3(1), dim3(128), PROBE_SMEM, nullptr, p);
  }
  PNP_CUDA(e);
  PNP_CUDA(cudaDeviceSynchronize());
  long long h[4];
  PNP_CUDA(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost));
  cudaFree(out);
  cycles_out_host[0] = h[0];
  cycles_out_host[1] = h[1];
  return 0;
}
