// One whole ConvNeXt Block (model.py:160-167) of the last stage (C = 256, hidden 512) per launch, 128 tokens per CTA:
//     out = x + (gamma W2) gelu( W1 LN( dwconv7(x) ) + b1 ) + gamma b2
// Replaces dwconv_ln_kernel<256> + gemm_tc2<256, BF16+GELU> + gemm_tc2<128, F32, RESID> (three launches, two bf16 round
// trips) for the three Blocks of stage 6.  Same skeleton as ffn_fused.cuh with a depthwise-conv + LayerNorm prologue:
//
//   warp 0      TMA producer: W1 [512, 256] and gamma*W2 [256, 512] streamed per hidden chunk through a ring of 32 KB stages
//   warp 1      tcgen05.mma issuer: per chunk of 64 hidden units MMA1 128x64x256 -> D1[c & 1], MMA2 128x256x64 -> D2 (+=)
//   warps 2-9   (a) dwconv7 + LN, one warp per token (register ring over rows) -> bf16 A operand (4 k-blocks)
//               (b) per chunk: bias + gelu(tanh) out of TMEM -> bf16 h chunk (one k-block of MMA2)
//               (c) D2 + gamma b2 staged through shared memory, coalesced out = stage + x
#pragma once
#include "ffn_fused.cuh"

namespace a2m {

constexpr int B6_CWARPS = 8;
constexpr int B6_CTHREADS = B6_CWARPS * 32;
constexpr int B6_THREADS = 64 + B6_CTHREADS;
constexpr int B6_C = 256, B6_H = 512, B6_CH = 64, B6_NCH = B6_H / B6_CH;   // 8 chunks
constexpr int B6_AUX_BYTES = (B6_H + B6_C) * 4 + 256;                       // b1, b2g, barriers, tmem slot
constexpr size_t B6_SMEM = 1024 + FF_MAIN_BYTES + B6_AUX_BYTES;

// tmW1: W1 [512, 256] bf16, box {64, 64};  tmW2: gamma*W2 [256, 512] bf16, box {64, 256}.
// params (fp32): dw[7][256] | dwb[256] | lnw[256] | lnb[256] | b1[512] | b2g[256]   (BigBlockW::fused)
__global__ void __launch_bounds__(B6_THREADS, 1)
block256_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const float* X, float* Y,
                      int L, int M, const float* __restrict__ params) {
  using RM = RowMap<B6_C>;
  constexpr int PER = RM::PER;   // 8 channels per lane
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sH = sA + FF_A_BYTES;
  uint8_t* sW = sH + 2 * FF_H_BYTES;
  float* sStage = reinterpret_cast<float*>(smem);   // aliases everything above once the last MMA has completed
  float* sB1 = reinterpret_cast<float*>(smem + FF_MAIN_BYTES);
  float* sB2 = sB1 + B6_H;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB2 + B6_C);
  uint64_t* bar_empty = bar_full + FF_NST;
  uint64_t* bar_a = bar_empty + FF_NST;
  uint64_t* bar_d1full = bar_a + 1;         // [2]
  uint64_t* bar_d1free = bar_d1full + 2;    // [2]
  uint64_t* bar_hfull = bar_d1free + 2;     // [2]
  uint64_t* bar_hfree = bar_hfull + 2;      // [2]
  uint64_t* bar_done = bar_hfree + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile0 = blockIdx.x * FF_ROWS;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int s = 0; s < FF_NST; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(bar_a, B6_CTHREADS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_d1full[i], 1);
      mbar_init(&bar_d1free[i], B6_CTHREADS);
      mbar_init(&bar_hfull[i], B6_CTHREADS);
      mbar_init(&bar_hfree[i], 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  copy_const_to_smem<(B6_H + B6_C) / 4, B6_THREADS>(sB1, params + 10 * B6_C, threadIdx.x);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d2 = tmem_base;             // columns 0..255
  const uint32_t tmem_d1 = tmem_base + 256;       // two accumulators of 64 columns

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer (constants: no pdl_wait needed)
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      auto load_w1 = [&](int c) {
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], FF_STAGE);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sW + s * FF_STAGE + kb * 8192, &tmW1, &bar_full[s], kb * 64, c * B6_CH);
        if (++s == FF_NST) { s = 0; ph ^= 1; }
      };
      auto load_w2 = [&](int c) {
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], FF_STAGE);
        tma_load_2d(sW + s * FF_STAGE, &tmW2, &bar_full[s], c * B6_CH, 0);
        if (++s == FF_NST) { s = 0; ph ^= 1; }
      };
      for (int c = 0; c <= B6_NCH; ++c) {
        if (c < B6_NCH) load_w1(c);
        if (c >= 1) load_w2(c - 1);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(128, B6_CH);
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, B6_C);
      uint32_t s = 0, ph = 0;
      mbar_wait(bar_a, 0);
      tc_fence_after();
      for (int c = 0; c <= B6_NCH; ++c) {
        if (c < B6_NCH) {
          mbar_wait(&bar_d1free[c & 1], ((c >> 1) & 1) ^ 1);
          mbar_wait(&bar_full[s], ph);
          tc_fence_after();
          const uint32_t d1 = tmem_d1 + (c & 1) * B6_CH;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint64_t da = umma_desc_sw128(smem_u32(sA + kb * (FF_ROWS * 128)));
            const uint64_t db = umma_desc_sw128(smem_u32(sW + s * FF_STAGE + kb * 8192));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d1, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(db, k * 32), idesc1, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_empty[s]);
          umma_commit(&bar_d1full[c & 1]);
          if (++s == FF_NST) { s = 0; ph ^= 1; }
        }
        if (c >= 1) {
          const int cc = c - 1;
          mbar_wait(&bar_hfull[cc & 1], (cc >> 1) & 1);
          mbar_wait(&bar_full[s], ph);
          tc_fence_after();
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
    const int cw = warp - 2;            // 0..7
    const int quad = warp & 3;          // TMEM lane quadrant
    const int chalf = cw >> 2;          // column half
    const int row = quad * 32 + lane;
    const uint32_t t_row = static_cast<uint32_t>(quad * 32) << 16;

    // (a) depthwise k7 + LayerNorm -> A operand: 16 tokens per warp in four passes of 4 (register ring over 10 rows)
    {
      float w[7][PER], bias[PER], lw[PER], lb[PER];
#pragma unroll
      for (int t = 0; t < 7; ++t) RM::load(params + t * B6_C, lane, w[t]);
      RM::load(params + 7 * B6_C, lane, bias);
      RM::load(params + 8 * B6_C, lane, lw);
      RM::load(params + 9 * B6_C, lane, lb);
      pdl_wait();  // weights / parameters above are constants; x is produced by the previous kernel
      constexpr int TPP = 4;   // 10 rows x 8 channels of x + 7 x 8 taps in registers
#pragma unroll 1
      for (int pass = 0; pass < 4; ++pass) {
        const int r0 = (cw * 4 + pass) * TPP;
        float rows[TPP + 6][PER];
#pragma unroll
        for (int i = 0; i < TPP + 6; ++i) {
          const int g = tile0 + r0 - 3 + i;
          if (g >= 0 && g < M) {
            RM::load(X + static_cast<size_t>(g) * B6_C, lane, rows[i]);
          } else {
#pragma unroll
            for (int j = 0; j < PER; ++j) rows[i][j] = 0.f;
          }
        }
        float y[TPP][PER];
        const int l_first = (tile0 + r0) % L;   // one runtime modulo per pass instead of one per token
#pragma unroll
        for (int i = 0; i < TPP; ++i) {
          const int l = l_first + i - ((l_first + i >= L) ? L : 0);
#pragma unroll
          for (int j = 0; j < PER; ++j) y[i][j] = bias[j];
#pragma unroll
          for (int t = 0; t < 7; ++t) {
            const int ll = l + t - 3;
            if (ll >= 0 && ll < L) {  // zero "SAME" padding at the window boundary (warp-uniform)
#pragma unroll
              for (int j = 0; j < PER; ++j) y[i][j] = fmaf(w[t][j], rows[i + t][j], y[i][j]);
            }
          }
        }
        // LayerNorm statistics of the 4 tokens jointly
        float st[TPP], mean[TPP];
#pragma unroll
        for (int i = 0; i < TPP; ++i) {
          float a = 0.f;
#pragma unroll
          for (int j = 0; j < PER; ++j) a += y[i][j];
          st[i] = a;
        }
        warp_sum_all<TPP>(st, lane);
#pragma unroll
        for (int i = 0; i < TPP; ++i) {
          mean[i] = st[i] * (1.0f / B6_C);
          float a = 0.f;
#pragma unroll
          for (int j = 0; j < PER; ++j) a += (y[i][j] - mean[i]) * (y[i][j] - mean[i]);
          st[i] = a;
        }
        warp_sum_all<TPP>(st, lane);
#pragma unroll
        for (int i = 0; i < TPP; ++i) {
          const int r = r0 + i;
          const float inv = rsqrtf(st[i] * (1.0f / B6_C) + kLnEps);
          const bool live = tile0 + r < M;
#pragma unroll
          for (int j = 0; j < PER; ++j) y[i][j] = live ? (y[i][j] - mean[i]) * inv * lw[j] + lb[j] : 0.f;
#pragma unroll
          for (int g = 0; g < RM::G; ++g) {
            const int col = RM::chan(lane, g);
            uint2 q;
            q.x = pack_bf16x2(y[i][4 * g], y[i][4 * g + 1]);
            q.y = pack_bf16x2(y[i][4 * g + 2], y[i][4 * g + 3]);
            *reinterpret_cast<uint2*>(sA + (col >> 6) * (FF_ROWS * 128) + sw128_offset(r, col & 63)) = q;
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_a);
    }

    // (b) bias + GELU of every hidden chunk: this thread owns 32 of the chunk's 64 hidden units of its row
#pragma unroll 1
    for (int c = 0; c < B6_NCH; ++c) {
      mbar_wait(&bar_d1full[c & 1], (c >> 1) & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_x32(tmem_d1 + (c & 1) * B6_CH + t_row + chalf * 32, r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bar_d1free[c & 1]);
      const float4* bg = reinterpret_cast<const float4*>(sB1 + c * B6_CH + chalf * 32);
      uint32_t packed[16];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 bv = bg[q];
        packed[2 * q] = pack_bf16x2(gelu_tanh_fast(__uint_as_float(r[4 * q]) + bv.x), gelu_tanh_fast(__uint_as_float(r[4 * q + 1]) + bv.y));
        packed[2 * q + 1] = pack_bf16x2(gelu_tanh_fast(__uint_as_float(r[4 * q + 2]) + bv.z), gelu_tanh_fast(__uint_as_float(r[4 * q + 3]) + bv.w));
      }
      mbar_wait(&bar_hfree[c & 1], ((c >> 1) & 1) ^ 1);   // MMA2 of chunk c-2 has finished reading this buffer
      uint8_t* hb = sH + (c & 1) * FF_H_BYTES;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(hb + sw128_offset(row, chalf * 32 + 8 * q)) =
            make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
      fence_proxy_async_smem();
      mbar_arrive(&bar_hfull[c & 1]);
    }

    // (c) D2 + gamma b2 -> staging, then coalesced out = stage + x
    mbar_wait(bar_done, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int col0 = chalf * 128 + c * 32;
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
    named_bar_sync(1, B6_CTHREADS);
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      float xv[8][PER];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = cw + (pass * 8 + i) * B6_CWARPS;
        if (tile0 + r < M) RM::load(X + static_cast<size_t>(tile0 + r) * B6_C, lane, xv[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = cw + (pass * 8 + i) * B6_CWARPS;
        if (tile0 + r < M) {
          float sv[PER];
          RM::load(sStage + r * FF_STAGE_STRIDE, lane, sv);
#pragma unroll
          for (int j = 0; j < PER; ++j) sv[j] += xv[i][j];
          RM::store_f32(Y + static_cast<size_t>(tile0 + r) * B6_C, lane, sv);
        }
      }
    }
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
