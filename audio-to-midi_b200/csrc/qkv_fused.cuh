// attention_norm + the folded q | k | v projection + RoPE of one TransformerLayer (model.py:539 -> SelfAttention
// model.py:353-362, rope.py:25-53) per launch, on one tile of 128 tokens per CTA:
//     qkv[:, 0:256]   = RoPE( LN(x) Wq^T )            (4 heads x 64)
//     qkv[:, 256:512] = RoPE( LN(x) (Wk Wc)^T )       (compressed-kv product folded at load, a2m_api.cu pack_weights)
//     qkv[:, 512:768] =        LN(x) (Wv Wc)^T
// Replaces ln_rows_kernel + gemm_tc2<128, ROPE>: the normalised activations never leave shared memory.
//
//   warp 0      TMA producer: W [768, 256] streamed as 24 stages of [128 rows x 64 k] (16 KB) through a 4-deep ring
//   warp 1      TMEM allocator + tcgen05.mma issuer: 3 column chunks of 256, accumulators ping-pong in 2 x 256 columns
//   warps 2-17  LayerNorm of the tile (one warp per row) -> bf16 A operand; then per chunk, in two halves of 32 columns per
//               head: TMEM -> RoPE -> bf16 into a 64B-swizzled staging tile -> TMA store (per-thread global stores of one
//               row each cost 32 LSU wavefronts per instruction and made the first version of this kernel store-bound)
// RoPE uses the absolute frame index inside the window (row % 256), as gemm_tc2's G2_ROPE epilogue does (see
// attention.cuh for why this equals the reference's per-local-window positions).
#pragma once
#include "ffn_fused.cuh"
#include "gemm_pair.cuh"

namespace a2m {

constexpr int QF_THREADS = FF_THREADS;          // 2 + 16 warps
constexpr int QF_N = 768;
constexpr int QF_NCHUNK = QF_N / 256;            // 3
// The weight stream paces the MMAs (in-kernel timeline, tools/ffn_timeline.py): 64 KB of ring is what the shared-memory budget
// leaves, and with two 32 KB stages the MMA holds half of it while it works, so only one stage is in flight against ~1000
// cycles of L2 latency.  Stages of 128 W rows (16 KB, UMMA N = 128 into one half of the chunk's accumulator) keep three in
// flight: 13.9 -> 13.6 us.  (Starting the CTAs at different column chunks, so that they do not all pull the same weight rows
// at the same moment, was measured SLOWER: simultaneous requests for the same lines are merged in L2.)
constexpr int QF_SROWS = 128;                    // W rows per ring stage
constexpr int QF_STAGE = QF_SROWS * 64 * 2;      // 16 KB
constexpr int QF_NST = 64 * 1024 / QF_STAGE;     // 4
constexpr int QF_NSUB = 256 / QF_SROWS;          // stages per (chunk, k-block)
constexpr int QF_ROPE_BYTES = 2 * FF_ROWS * 32 * 4;               // cos then sin, 128 positions x 32 pairs (float4 index XOR row & 7)
constexpr int QF_OUT_TILE = FF_ROWS * 64;                         // staging tile: 128 rows x 32 bf16, 64B swizzle
constexpr int QF_OUT_BYTES = 2 * 4 * QF_OUT_TILE;                 // one tile per head of the chunk, double buffered
constexpr int QF_MAIN_BYTES = FF_A_BYTES + QF_NST * QF_STAGE;     // 128 KB
constexpr size_t QF_SMEM = 1024 + QF_MAIN_BYTES + QF_OUT_BYTES + QF_ROPE_BYTES + 256;

// tmW: Wqkv [768, 256] bf16, box {64, QF_SROWS}.  tmO: out [M, 768] bf16, box {32, 128}, 64B swizzle.  X: fp32 [M, 256].
// rope_cos / rope_sin: [>= 256, 32].
__global__ void __launch_bounds__(QF_THREADS, 1)
qkv_fused_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const float* X, int M,
                 const float* __restrict__ lnw, const float* __restrict__ lnb, const float* __restrict__ rope_cos,
                 const float* __restrict__ rope_sin, int rows_per_window) {
  using RM = RowMap<FF_D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sW = sA + FF_A_BYTES;
  uint8_t* sOut = smem + QF_MAIN_BYTES;
  float* sCos = reinterpret_cast<float*>(sOut + QF_OUT_BYTES);
  float* sSin = sCos + FF_ROWS * 32;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sSin + FF_ROWS * 32);
  uint64_t* bar_empty = bar_full + QF_NST;
  uint64_t* bar_a = bar_empty + QF_NST;
  uint64_t* bar_dfull = bar_a + 1;      // [2]
  uint64_t* bar_dfree = bar_dfull + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_dfree + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile0 = blockIdx.x * FF_ROWS;
  const int pos0 = tile0 % rows_per_window;    // a tile never straddles windows (rows_per_window is a multiple of 128)

  pdl_launch_dependents();
  if (threadIdx.x == 0) FF_STAMP(80);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < QF_NST; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(bar_a, FF_CTHREADS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_dfull[i], 1);
      mbar_init(&bar_dfree[i], FF_CTHREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // RoPE rows pos0 .. pos0+127 (constants): 128 x 8 float4 each for cos and sin; all loads of a thread before its stores
  {
    constexpr int NV = 2 * FF_ROWS * 8, PER = (NV + QF_THREADS - 1) / QF_THREADS;
    float4 v[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = threadIdx.x + k * QF_THREADS;
      const int which = i >> 10, r = (i >> 3) & 127, q = i & 7;
      if (i < NV) v[k] = __ldg(reinterpret_cast<const float4*>((which ? rope_sin : rope_cos) + (pos0 + r) * 32) + q);
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = threadIdx.x + k * QF_THREADS;
      const int which = i >> 10, r = (i >> 3) & 127, q = i & 7;
      if (i < NV) *reinterpret_cast<float4*>((which ? sSin : sCos) + r * 32 + 4 * (q ^ (r & 7))) = v[k];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) FF_STAMP(81);

  if (warp == 0) {
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      for (int n = 0; n < QF_NCHUNK; ++n)
        for (int sub = 0; sub < QF_NSUB; ++sub)
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(&bar_empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&bar_full[s], QF_STAGE);
            tma_load_2d(sW + s * QF_STAGE, &tmW, &bar_full[s], kb * 64, n * 256 + sub * QF_SROWS);
            if (++s == QF_NST) { s = 0; ph ^= 1; }
          }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, QF_SROWS);
      uint32_t s = 0, ph = 0;
      mbar_wait(bar_a, 0);
      tc_fence_after();
      FF_STAMP(83);
      for (int n = 0; n < QF_NCHUNK; ++n) {
        mbar_wait(&bar_dfree[n & 1], ((n >> 1) & 1) ^ 1);
        tc_fence_after();
        FF_STAMP(84 + n * 6);
        for (int sub = 0; sub < QF_NSUB; ++sub) {
          const uint32_t d = tmem_base + (n & 1) * 256 + sub * QF_SROWS;
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(&bar_full[s], ph);
            tc_fence_after();
            if (sub == QF_NSUB - 1) FF_STAMP(84 + n * 6 + 1 + kb);
            const uint64_t da = umma_desc_sw128(smem_u32(sA + kb * (FF_ROWS * 128)));
            const uint64_t db = umma_desc_sw128(smem_u32(sW + s * QF_STAGE));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(db, k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&bar_empty[s]);
            if (++s == QF_NST) { s = 0; ph ^= 1; }
          }
        }
        umma_commit(&bar_dfull[n & 1]);
      }
    }
  } else {
    const int cw = warp - 2;
    const int quad = warp & 3;
    const int cq = cw >> 2;             // head index inside the chunk
    const int row = quad * 32 + lane;
    const uint32_t t_row = static_cast<uint32_t>(quad * 32) << 16;

    ff_layer_norm_to_operand(X, M, tile0, lnw, lnb, sA, cw, lane);
    fence_proxy_async_smem();
    mbar_arrive(bar_a);
    if (threadIdx.x == 64) FF_STAMP(82);

    const uint32_t rsw = static_cast<uint32_t>(row >> 1) & 3u;   // 64B swizzle: 16-byte chunk index ^= (row / 2) % 4
    const bool storer = threadIdx.x == 64;
#pragma unroll 1
    for (int n = 0; n < QF_NCHUNK; ++n) {
      mbar_wait(&bar_dfull[n & 1], (n >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 64) FF_STAMP(104 + n * 2);
      const uint32_t d = tmem_base + (n & 1) * 256 + t_row + cq * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld_x32(d + half * 32, r);
        tmem_ld_wait();
        if (half == 1) {   // this thread's part of the accumulator is in registers: hand the buffer back
          tc_fence_before();
          mbar_arrive(&bar_dfree[n & 1]);
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (n < 2) {   // q and k: rotate pairs (2i, 2i+1), i = half * 16 + 4 q + t  (rope.py:43-52)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int q4 = (half * 4 + q) ^ (row & 7);
            const float4 cc = *reinterpret_cast<const float4*>(sCos + row * 32 + 4 * q4);
            const float4 ss = *reinterpret_cast<const float4*>(sSin + row * 32 + 4 * q4);
            const float c4[4] = {cc.x, cc.y, cc.z, cc.w}, s4[4] = {ss.x, ss.y, ss.z, ss.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float x1 = v[8 * q + 2 * t], x2 = v[8 * q + 2 * t + 1];
              v[8 * q + 2 * t] = x1 * c4[t] - x2 * s4[t];
              v[8 * q + 2 * t + 1] = x1 * s4[t] + x2 * c4[t];
            }
          }
        }
        // staging tiles are double buffered by half: the stores of the previous half (other buffer) are waited for just
        // before this half's barrier, so the buffer the NEXT half overwrites is known to be drained by then
        uint8_t* sbuf = sOut + half * (4 * QF_OUT_TILE);
        uint8_t* stile = sbuf + cq * QF_OUT_TILE;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(v[8 * q], v[8 * q + 1]);
          o.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
          o.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
          o.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
          *reinterpret_cast<uint4*>(stile + row * 64 + ((static_cast<uint32_t>(q) ^ rsw) << 4)) = o;
        }
        fence_proxy_async_smem();
        if (storer) bulk_wait_read<0>();
        named_bar_sync(1, FF_CTHREADS);
        if (storer) {
#pragma unroll
          for (int hd = 0; hd < 4; ++hd) tma_store_2d(&tmO, sbuf + hd * QF_OUT_TILE, n * 256 + hd * 64 + half * 32, tile0);
          bulk_commit();
        }
      }
      if (threadIdx.x == 64) FF_STAMP(104 + n * 2 + 1);
    }
    if (storer) bulk_wait_all<0>();   // the stores must have landed before the grid counts as complete
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace a2m
