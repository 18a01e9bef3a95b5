// gemm_pair_kernel: the CTA-pair (tcgen05 cta_group::2) recipe, proven on a plain GEMM before it is applied to the fused
// transformer kernels (the B operand split over two SMs halves both the L2 weight stream and the shared-memory
// operand reads that bound qkv_fused_kernel / postattn_fused_kernel).
//
//     D[M, N] (fp32) = A[M, K] (bf16, K-major) x W[N, K]^T (bf16, K-major),   N in {64, 128, 256},  K % 64 == 0
//
// One cluster of two CTAs per 256 rows of A.  CTA r of the pair
//   * TMA-loads ITS 128 rows of A and ITS half of W (rows r N/2 .. r N/2 + N/2 - 1) into its own shared memory, at the same
//     offsets in both CTAs (the pair's MMA addresses both shared memories with one descriptor);
//   * owns the accumulator of its 128 rows in its own tensor memory (128 lanes x N columns);
//   * the peer (rank 1) tells the leader that a stage has landed with a remote mbarrier arrive (mapa + arrive.release.cluster),
//     issued by its otherwise idle MMA warp;
//   * only the leader (rank 0) issues tcgen05.mma.cta_group::2 (M = 256) and commits with a multicast arrive, so both CTAs see
//     "stage consumed" and "accumulator complete" on their own barriers.
#pragma once
#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace a2m {

constexpr int GP_THREADS = 192;   // warp 0 TMA, warp 1 MMA (leader only), warps 2-5 epilogue
constexpr int GP_STAGES = 4;

template <int BN>
constexpr size_t gemm_pair_smem_bytes() {
  return 1024 + GP_STAGES * (128 * 64 * 2 + (BN / 2) * 64 * 2) + 256;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_ptr` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(const void* local_smem_ptr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst) {  // whole warp, in BOTH CTAs of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols));
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all previously issued MMAs of the pair have completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

// tmA: A [M, K] bf16 box {64, 128};  tmB: W [N, K] bf16 box {64, BN / 2}.  D: fp32 [M, BN] row-major (ldd).
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GP_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ D, int ldd, int M, int K) {
  constexpr int A_BYTES = 128 * 64 * 2;
  constexpr int B_BYTES = (BN / 2) * 64 * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = sA + GP_STAGES * A_BYTES;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB + GP_STAGES * B_BYTES);   // own TMA bytes landed
  uint64_t* bar_peer = bar_full + GP_STAGES;                                    // leader: the peer's stage has landed
  uint64_t* bar_empty = bar_peer + GP_STAGES;                                   // stage consumed by the pair's MMAs
  uint64_t* bar_done = bar_empty + GP_STAGES;                                   // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int row0 = pair * 256 + static_cast<int>(rank) * 128;   // this CTA's rows of A / D
  const int num_kb = K / 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < GP_STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_peer[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_cluster(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], A_BYTES + B_BYTES);
        tma_load_2d(sA + s * A_BYTES, &tmA, &bar_full[s], kb * 64, row0);
        tma_load_2d(sB + s * B_BYTES, &tmB, &bar_full[s], kb * 64, static_cast<int>(rank) * (BN / 2));
        if (++s == GP_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
      uint32_t s = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&bar_full[s], ph);
        mbar_wait_cluster(&bar_peer[s], ph);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem_u32(sA + s * A_BYTES));
        const uint64_t db = umma_desc_sw128(smem_u32(sB + s * B_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_pair(tmem_base, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(db, k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit_pair(&bar_empty[s]);
        if (++s == GP_STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit_pair(bar_done);
    } else if (rank == 1 && elect_one()) {
      // the peer's MMA warp has nothing to issue: it tells the leader when each of this CTA's stages has landed
      uint32_t s = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&bar_full[s], ph);
        mbar_arrive_remote(mapa_shared(&bar_peer[s], 0));
        if (++s == GP_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    mbar_wait_cluster(bar_done, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    float* drow = D + static_cast<size_t>(row0 + r) * ldd;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_x32(taddr + c * 32, v);
      tmem_ld_wait();
      if (row0 + r < M) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          reinterpret_cast<float4*>(drow + c * 32)[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                     __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA may free its tensor memory / exit while the pair's MMAs or remote arrives can touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<BN>(tmem_base);
  }
}

}  // namespace a2m
