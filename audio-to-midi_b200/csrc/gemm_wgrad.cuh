// gemm_wgrad_kernel: weight gradients of every tensor-core layer on tcgen05,
//     dW[N_out, K_out] += dY[T, N_out]^T . X[T, K_out]          (bf16 operands, fp32 accumulate)
// i.e. the `eqx.filter_value_and_grad` contribution of a Linear / 1x1 Conv layer (train.py:50; the forward
// call sites are listed in gemm_tc.cuh).  The reduction runs over TOKENS, the slow axis of both row-major
// operands, so both tiles are fed to the tensor core as MN-major operands: a pipeline stage holds 64 tokens of
// dY (128 columns = two 64-column TMA boxes) and of X (BN columns = BN/64 boxes), each box [64 tokens][128 B]
// 128B-swizzled exactly as the TMA unit wrote it; the UMMA descriptors name the box pitch as the leading byte
// offset and set the transpose bits of A and B.
//
// Work split: grid = output tiles (128 x BN of dW) x token splits; each CTA reduces its token range into one
// TMEM accumulator and adds it to dW with vectorised fp32 reductions (red.global.add.v4.f32), which also gives
// the accumulate-into-gradient-buffer semantics of minibatch accumulation (train.py:283-293).
// Rows of dY/X beyond T and columns beyond N_out / K_out are zero-filled by the TMA unit.
#pragma once
#include "ptx.cuh"

namespace a2m {

constexpr int WG_TOK = 64;        // tokens per pipeline stage
constexpr int WG_THREADS = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int WG_BOX = WG_TOK * 128;  // bytes of one [64 tokens x 64 columns] bf16 box

template <int BN>
__host__ __device__ constexpr int wg_stages() { return BN == 256 ? 4 : 6; }
template <int BN>
constexpr size_t wgrad_smem_bytes() {
  return 1024 + wg_stages<BN>() * (2 * WG_BOX + (BN / 64) * WG_BOX) + 256;
}

// MN-major operand tile: boxes of [rows = K (tokens)][64 MN elements = 128 B], 128B swizzle; consecutive 64-wide
// MN chunks are `lbo_bytes` apart, 8-row groups along K are 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// both operands MN-major (transpose bits 15 and 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_abmn(uint32_t m, uint32_t n) {
  return umma_idesc_bf16(m, n) | (1u << 15) | (1u << 16);
}
// A MN-major, B K-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16_amn(uint32_t m, uint32_t n) {
  return umma_idesc_bf16(m, n) | (1u << 15);
}

// TMA reduction smem -> global: adds a 128B-swizzled [rows x 32 fp32] box into the fp32 tensor behind `m` (element-wise
// add performed by the L2, one full line per request instead of one 16-byte atomic per thread); rows / columns outside
// the tensor are clipped.
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

// tmA: dY [T, N_out] bf16, box {64 cols, 64 rows};  tmB: X [T, K_out] bf16, box {64 cols, 64 rows};
// tmD: dW [N_out, K_out] fp32, box {32 cols, 128 rows} (128B swizzle).
// bias_out (optional): fp32 [N_out] += column sums of dY (the bias gradient that goes with this weight gradient),
// computed by the otherwise idle epilogue warps from the staged dY tiles of the CTAs that own column tile 0.
template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmD, float* __restrict__ bias_out, int n_out, int k_out, int tokens, int splits) {
  constexpr int STAGES = wg_stages<BN>();
  constexpr int A_BYTES = 2 * WG_BOX;
  constexpr int B_BYTES = (BN / 64) * WG_BOX;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  static_assert(STAGES * (A_BYTES + B_BYTES) >= (BN / 32) * 128 * 128, "epilogue staging re-uses the pipeline buffers");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * A_BYTES;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* bar_empty = bar_full + STAGES;
  uint64_t* bar_done = bar_empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kt = (k_out + BN - 1) / BN;
  const int tile = static_cast<int>(blockIdx.x) / splits, split = static_cast<int>(blockIdx.x) % splits;
  const int n_blk = tile / num_kt, k_blk = tile % num_kt;
  const int nb = (tokens + WG_TOK - 1) / WG_TOK;
  const int tb0 = static_cast<int>(static_cast<long long>(nb) * split / splits);
  const int tb1 = static_cast<int>(static_cast<long long>(nb) * (split + 1) / splits);
  const bool do_bias = bias_out != nullptr && k_blk == 0;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], do_bias ? 5 : 1);   // MMA commit (+ one lane of each epilogue warp)
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (tb1 <= tb0) {  // nothing to reduce (more splits than token blocks)
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
    return;
  }

  if (warp == 0) {
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      for (int tb = tb0; tb < tb1; ++tb) {
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], A_BYTES + B_BYTES);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          tma_load_2d(sA + s * A_BYTES + c * WG_BOX, &tmA, &bar_full[s], n_blk * 128 + c * 64, tb * WG_TOK);
#pragma unroll
        for (int c = 0; c < BN / 64; ++c)
          tma_load_2d(sB + s * B_BYTES + c * WG_BOX, &tmB, &bar_full[s], k_blk * BN + c * 64, tb * WG_TOK);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16_abmn(128, BN);
      uint32_t s = 0, ph = 0;
      for (int tb = tb0; tb < tb1; ++tb) {
        mbar_wait(&bar_full[s], ph);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128_mn(smem_u32(sA + s * A_BYTES), WG_BOX);
        const uint64_t db = umma_desc_sw128_mn(smem_u32(sB + s * B_BYTES), WG_BOX);
#pragma unroll
        for (int k = 0; k < WG_TOK / 16; ++k)   // 16 tokens = 16 rows of 128 B
          umma_bf16(tmem_base, umma_desc_advance_k(da, k * 2048), umma_desc_advance_k(db, k * 2048), idesc,
                    (tb != tb0 || k != 0) ? 1u : 0u);
        umma_commit(&bar_empty[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(bar_done);
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;          // accumulator row == column of the dY tile
    if (do_bias) {
      // column sums of the staged dY tiles: element (tok, r) of chunk r / 64 in the 128B-swizzled box
      float acc = 0.f;
      uint32_t s = 0, ph = 0;
      const uint32_t cbase = (r >> 6) * WG_BOX + (r & 7) * 2, c16 = (r & 63) >> 3;
      for (int tb = tb0; tb < tb1; ++tb) {
        mbar_wait(&bar_full[s], ph);
        const uint8_t* base = sA + s * A_BYTES + cbase;
#pragma unroll 8
        for (int t = 0; t < WG_TOK; ++t)
          acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(base + t * 128 + ((c16 ^ (t & 7)) << 4)));
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      const int n = n_blk * 128 + r;
      if (n < n_out) atomicAdd(bias_out + n, acc);
    }
    mbar_wait(bar_done, 0);
    tc_fence_after();
    // bar_done only says the MMAs have consumed every stage; a slower epilogue warp may still be summing bias columns
    // out of the last stages, which the staging writes below overwrite
    if (do_bias) named_bar_sync(1, 128);
    // accumulator -> fp32 staging chunks (128 rows x 32 columns, swizzled) in the drained pipeline buffers -> TMA add
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t rsw = static_cast<uint32_t>(r & 7);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_x32(taddr + c * 32, v);
      tmem_ld_wait();
      uint8_t* srow = smem + c * (128 * 128) + r * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(srow + ((static_cast<uint32_t>(q) ^ rsw) << 4)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (warp == 2 && lane == 0) {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c)
        if (k_blk * BN + c * 32 < k_out) tma_reduce_add_2d(&tmD, smem + c * (128 * 128), k_blk * BN + c * 32, n_blk * 128);
      bulk_commit();
      bulk_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace a2m
