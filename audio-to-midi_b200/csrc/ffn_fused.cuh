// One whole FeedForwardBlock sub-layer of a TransformerLayer (model.py:546-553 with FeedForwardBlock.__call__,
// model.py:226-238) per launch, on one tile of 128 tokens per CTA:
//     x <- x + W2 ( gelu(u[:512]) * u[512:] ) + b2,      u = W1 LN(x) + b1
// Everything between the first read of x and the final store stays on chip: the LayerNorm output, the 1024-wide
// pre-activation u (TMEM only) and the 512-wide gated activation h (shared memory only) never reach L2 / HBM.
// Replaces ln_rows_kernel + gemm_tc2<256, GLU> + gemm_tc2<128, F32, RESID> (three launches, two bf16 round trips).
//
//   warp 0      TMA producer: streams W1 / W2 through a ring of 32 KB stages (768 KB per CTA, L2-resident)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-17  (a) LayerNorm of the tile, one warp per row, written as the bf16 A operand (4 k-blocks, 128B swizzle)
//               (b) per hidden chunk: bias + gelu(tanh) * gate out of TMEM -> bf16 h chunk (one k-block of MMA2)
//               (c) final: D2 + b2 staged through shared memory, then coalesced x + stage -> x
//
// The hidden dimension is processed in 8 chunks of 64 units.  W1's rows are packed so that chunk c holds its 64 "gelu"
// rows followed by its 64 "gate" rows (model.py:233-234), i.e. MMA1 of a chunk is a 128 x 128 x 256 product into one of
// two TMEM accumulators (ping-pong), and MMA2 of a chunk is a 128 x 256 x 64 product accumulated over chunks into D2.
// Issue order MMA1(c+1), MMA2(c): the tensor pipe works on the next chunk's pre-activation while the CUDA cores gate
// the current one.
#pragma once
#include "cnn_kernels.cuh"
#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace a2m {

#ifdef A2M_FFN_TIMING
#define FF_STAMP(i) do { if (blockIdx.x == 0) g_ffn_timing[(i)] = clock64(); } while (0)
#else
#define FF_STAMP(i) do { } while (0)
#endif

constexpr int FF_CWARPS = 16;                    // compute warps: TMEM quadrant = warp & 3, column quarter = (warp - 2) >> 2
constexpr int FF_CTHREADS = FF_CWARPS * 32;
constexpr int FF_THREADS = 64 + FF_CTHREADS;
constexpr int FF_ROWS = 128;
constexpr int FF_D = 256;            // model width
constexpr int FF_F = 512;            // gated width (FeedForwardBlock intermediate_size)
constexpr int FF_CH = 64;            // hidden units per chunk
constexpr int FF_NCH = FF_F / FF_CH; // 8
constexpr int FF_STAGE = 32 * 1024;  // ring stage: W1 [128 rows x 128 k] or W2 [256 rows x 64 k]
constexpr int FF_NST = 3;
constexpr int FF_A_BYTES = FF_ROWS * FF_D * 2;        // 64 KB: 4 k-blocks of [128 x 64]
constexpr int FF_H_BYTES = FF_ROWS * FF_CH * 2;       // 16 KB per h chunk, double buffered
constexpr int FF_STAGE_STRIDE = FF_D + 4;             // floats per staged row (conflict-free float4 rows)
constexpr int FF_MAIN_BYTES = FF_A_BYTES + 2 * FF_H_BYTES + FF_NST * FF_STAGE;   // 192 KB
static_assert(FF_ROWS * FF_STAGE_STRIDE * 4 <= FF_MAIN_BYTES, "final staging re-uses the operand bytes");
constexpr int FF_AUX_BYTES = (2 * FF_F + FF_D) * 4 + 256;   // b1, b2, barriers, tmem slot
constexpr size_t FF_SMEM = 1024 + FF_MAIN_BYTES + FF_AUX_BYTES;

// LayerNorm of the 128-row tile, one warp per row (8 rows per compute warp, all their loads in flight at once), written
// as the bf16 A operand: 4 k-blocks of [128 rows x 64 channels], 128B-swizzled.  Rows beyond M are treated as zeros.
__device__ __forceinline__ void ff_layer_norm_to_operand(const float* X, int M, int tile0, const float* __restrict__ lnw,
                                                         const float* __restrict__ lnb, uint8_t* sA, int cw, int lane) {
  using RM = RowMap<FF_D>;
  pdl_wait();   // x is produced by the previous kernel
  if (threadIdx.x == 64) FF_STAMP(2);
  float xv[8][RM::PER];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = cw * 8 + i;
    if (tile0 + r < M) {
      RM::load(X + static_cast<size_t>(tile0 + r) * FF_D, lane, xv[i]);
    } else {
#pragma unroll
      for (int j = 0; j < RM::PER; ++j) xv[i][j] = 0.f;
    }
  }
  float lw[RM::PER], lb[RM::PER];
  RM::load(lnw, lane, lw);
  RM::load(lnb, lane, lb);
  // statistics of the 8 rows jointly (17 shuffles per reduction instead of 40 dependent ones)
  float st[8], mean[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < RM::PER; ++j) a += xv[i][j];
    st[i] = a;
  }
  warp_sum8_all(st, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mean[i] = st[i] * (1.0f / FF_D);
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < RM::PER; ++j) a += (xv[i][j] - mean[i]) * (xv[i][j] - mean[i]);
    st[i] = a;
  }
  warp_sum8_all(st, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = cw * 8 + i;
    const float inv = rsqrtf(st[i] * (1.0f / FF_D) + kLnEps);
#pragma unroll
    for (int j = 0; j < RM::PER; ++j) xv[i][j] = (xv[i][j] - mean[i]) * inv * lw[j] + lb[j];
#pragma unroll
    for (int g = 0; g < RM::G; ++g) {
      const int col = RM::chan(lane, g);
      uint2 q;
      q.x = pack_bf16x2(xv[i][4 * g], xv[i][4 * g + 1]);
      q.y = pack_bf16x2(xv[i][4 * g + 2], xv[i][4 * g + 3]);
      *reinterpret_cast<uint2*>(sA + (col >> 6) * (FF_ROWS * 128) + sw128_offset(r, col & 63)) = q;
    }
  }
}

// tmW1: packed W1 [1024, 256] bf16, box {64, 128};  tmW2: W2 [256, 512] bf16, box {64, 256}.
// X: fp32 [M, 256] residual stream, updated in place.  lnw / lnb: feed_forward_norm.  b1p: packed like W1's rows.
__global__ void __launch_bounds__(FF_THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, float* X, int M,
                 const float* __restrict__ lnw, const float* __restrict__ lnb, const float* __restrict__ b1p,
                 const float* __restrict__ b2) {
  using RM = RowMap<FF_D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sA = smem;
  uint8_t* sH = sA + FF_A_BYTES;
  uint8_t* sW = sH + 2 * FF_H_BYTES;
  float* sStage = reinterpret_cast<float*>(smem);   // aliases everything above once the last MMA has completed
  float* sB1 = reinterpret_cast<float*>(smem + FF_MAIN_BYTES);
  float* sB2 = sB1 + 2 * FF_F;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB2 + FF_D);
  uint64_t* bar_empty = bar_full + FF_NST;
  uint64_t* bar_a = bar_empty + FF_NST;     // A operand (LayerNorm output) ready
  uint64_t* bar_d1full = bar_a + 1;         // [2]
  uint64_t* bar_d1free = bar_d1full + 2;    // [2]
  uint64_t* bar_hfull = bar_d1free + 2;     // [2]
  uint64_t* bar_hfree = bar_hfull + 2;      // [2]
  uint64_t* bar_done = bar_hfree + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile0 = blockIdx.x * FF_ROWS;

  pdl_launch_dependents();
  if (threadIdx.x == 0) FF_STAMP(0);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int s = 0; s < FF_NST; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(bar_a, FF_CTHREADS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_d1full[i], 1);
      mbar_init(&bar_d1free[i], FF_CTHREADS);
      mbar_init(&bar_hfull[i], FF_CTHREADS);
      mbar_init(&bar_hfree[i], 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (threadIdx.x < (2 * FF_F + FF_D) / 4) {   // 320 float4: b1 (packed) then b2
    const int i = threadIdx.x;
    reinterpret_cast<float4*>(sB1)[i] = (i < 2 * FF_F / 4) ? __ldg(reinterpret_cast<const float4*>(b1p) + i)
                                                           : __ldg(reinterpret_cast<const float4*>(b2) + i - 2 * FF_F / 4);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d2 = tmem_base;             // columns 0..255
  const uint32_t tmem_d1 = tmem_base + 256;       // two accumulators of 128 columns
  if (threadIdx.x == 0) FF_STAMP(1);

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer (constants: no pdl_wait needed)
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      auto load_w1 = [&](int c, int half) {
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], FF_STAGE);
        tma_load_2d(sW + s * FF_STAGE, &tmW1, &bar_full[s], (2 * half) * 64, c * 128);
        tma_load_2d(sW + s * FF_STAGE + 16384, &tmW1, &bar_full[s], (2 * half + 1) * 64, c * 128);
        if (++s == FF_NST) { s = 0; ph ^= 1; }
      };
      auto load_w2 = [&](int c) {
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], FF_STAGE);
        tma_load_2d(sW + s * FF_STAGE, &tmW2, &bar_full[s], c * FF_CH, 0);
        if (++s == FF_NST) { s = 0; ph ^= 1; }
      };
      for (int c = 0; c <= FF_NCH; ++c) {
        if (c < FF_NCH) { load_w1(c, 0); load_w1(c, 1); }
        if (c >= 1) load_w2(c - 1);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(128, 128);
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, FF_D);
      uint32_t s = 0, ph = 0;
      mbar_wait(bar_a, 0);
      tc_fence_after();
      FF_STAMP(4);
      for (int c = 0; c <= FF_NCH; ++c) {
        if (c < FF_NCH) {
          mbar_wait(&bar_d1free[c & 1], ((c >> 1) & 1) ^ 1);
          tc_fence_after();
          FF_STAMP(8 + c * 4);
          const uint32_t d1 = tmem_d1 + (c & 1) * 128;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            mbar_wait(&bar_full[s], ph);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t da = umma_desc_sw128(smem_u32(sA + (2 * half + j) * (FF_ROWS * 128)));
              const uint64_t db = umma_desc_sw128(smem_u32(sW + s * FF_STAGE + j * 16384));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d1, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(db, k * 32), idesc1, (half | j | k) != 0 ? 1u : 0u);
            }
            umma_commit(&bar_empty[s]);
            if (++s == FF_NST) { s = 0; ph ^= 1; }
          }
          umma_commit(&bar_d1full[c & 1]);
          FF_STAMP(8 + c * 4 + 1);
        }
        if (c >= 1) {
          const int cc = c - 1;
          mbar_wait(&bar_hfull[cc & 1], (cc >> 1) & 1);
          FF_STAMP(8 + cc * 4 + 2);
          mbar_wait(&bar_full[s], ph);
          tc_fence_after();
          FF_STAMP(8 + cc * 4 + 3);
          const uint64_t da = umma_desc_sw128(smem_u32(sH + (cc & 1) * FF_H_BYTES));
          const uint64_t db = umma_desc_sw128(smem_u32(sW + s * FF_STAGE));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_d2, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(db, k * 32), idesc2, (cc | k) != 0 ? 1u : 0u);
          umma_commit(&bar_empty[s]);
          umma_commit(&bar_hfree[cc & 1]);
          if (++s == FF_NST) { s = 0; ph ^= 1; }
        }
      }
      umma_commit(bar_done);
    }
  } else {
    // ------------------------------------------------------------ compute warps
    const int cw = warp - 2;            // 0..15
    const int quad = warp & 3;          // TMEM lane quadrant this warp may read
    const int cq = cw >> 2;             // which quarter of the columns
    const int row = quad * 32 + lane;
    const uint32_t t_row = static_cast<uint32_t>(quad * 32) << 16;

    // (a) LayerNorm -> A operand
    ff_layer_norm_to_operand(X, M, tile0, lnw, lnb, sA, cw, lane);
    fence_proxy_async_smem();
    mbar_arrive(bar_a);
    if (threadIdx.x == 64) FF_STAMP(3);

    // (b) gate every hidden chunk: this thread owns 16 of the chunk's 64 hidden units of its row
#pragma unroll 1
    for (int c = 0; c < FF_NCH; ++c) {
      mbar_wait(&bar_d1full[c & 1], (c >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 64) FF_STAMP(48 + c * 3);
      const uint32_t d1 = tmem_d1 + (c & 1) * 128 + t_row;
      uint32_t r1[16], r2[16];
      tmem_ld_x16(d1 + cq * 16, r1);
      tmem_ld_x16(d1 + 64 + cq * 16, r2);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bar_d1free[c & 1]);
      if (threadIdx.x == 64) FF_STAMP(48 + c * 3 + 1);
      const float4* bg = reinterpret_cast<const float4*>(sB1 + c * 128 + cq * 16);   // gelu-row biases; gate rows 64 further
      uint32_t packed[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b1v = bg[q], b2v = bg[16 + q];
        const float h0 = gelu_tanh_fast(__uint_as_float(r1[4 * q]) + b1v.x) * (__uint_as_float(r2[4 * q]) + b2v.x);
        const float h1 = gelu_tanh_fast(__uint_as_float(r1[4 * q + 1]) + b1v.y) * (__uint_as_float(r2[4 * q + 1]) + b2v.y);
        const float h2 = gelu_tanh_fast(__uint_as_float(r1[4 * q + 2]) + b1v.z) * (__uint_as_float(r2[4 * q + 2]) + b2v.z);
        const float h3 = gelu_tanh_fast(__uint_as_float(r1[4 * q + 3]) + b1v.w) * (__uint_as_float(r2[4 * q + 3]) + b2v.w);
        packed[2 * q] = pack_bf16x2(h0, h1);
        packed[2 * q + 1] = pack_bf16x2(h2, h3);
      }
      mbar_wait(&bar_hfree[c & 1], ((c >> 1) & 1) ^ 1);   // MMA2 of chunk c-2 has finished reading this buffer
      uint8_t* hb = sH + (c & 1) * FF_H_BYTES;
#pragma unroll
      for (int q = 0; q < 2; ++q)
        *reinterpret_cast<uint4*>(hb + sw128_offset(row, cq * 16 + 8 * q)) =
            make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
      fence_proxy_async_smem();
      mbar_arrive(&bar_hfull[c & 1]);
      if (threadIdx.x == 64) FF_STAMP(48 + c * 3 + 2);
    }

    // (c) D2 + b2 -> staging (all operand bytes are dead once bar_done fires), then coalesced x += stage
    mbar_wait(bar_done, 0);
    tc_fence_after();
    if (threadIdx.x == 64) FF_STAMP(5);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col0 = cq * 64 + c * 32;
      uint32_t r[32];
      tmem_ld_x32(tmem_d2 + t_row + col0, r);
      tmem_ld_wait();
      float* srow = sStage + row * FF_STAGE_STRIDE + col0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        reinterpret_cast<float4*>(srow)[q] =
            make_float4(__uint_as_float(r[4 * q]) + sB2[col0 + 4 * q], __uint_as_float(r[4 * q + 1]) + sB2[col0 + 4 * q + 1],
                        __uint_as_float(r[4 * q + 2]) + sB2[col0 + 4 * q + 2], __uint_as_float(r[4 * q + 3]) + sB2[col0 + 4 * q + 3]);
    }
    named_bar_sync(1, FF_CTHREADS);
    if (threadIdx.x == 64) FF_STAMP(6);
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      float xv[4][RM::PER];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = cw + (pass * 4 + i) * FF_CWARPS;
        if (tile0 + r < M) RM::load(X + static_cast<size_t>(tile0 + r) * FF_D, lane, xv[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = cw + (pass * 4 + i) * FF_CWARPS;
        if (tile0 + r < M) {
          float sv[RM::PER];
          RM::load(sStage + r * FF_STAGE_STRIDE, lane, sv);
#pragma unroll
          for (int j = 0; j < RM::PER; ++j) sv[j] += xv[i][j];
          RM::store_f32(X + static_cast<size_t>(tile0 + r) * FF_D, lane, sv);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) FF_STAMP(7);
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace a2m
