// One whole ConvNeXt Block (model.py:160-167) per launch for C in {64, 128}:
//     out = x + gamma * pw2( gelu( pw1( LN( dwconv7(x) ) ) ) )
// on one tile of 128 tokens per CTA, everything between the first read of x and the final store on chip.
// Sized so that TWO CTAs fit on an SM (<= 100 KB smem, 256 TMEM columns, 256 threads): a tile is a chain of
// dependent phases (loads -> LN -> MMA -> GELU -> MMA -> store), and the second CTA fills the bubbles.
//
//   phase 1  CUDA cores   depthwise k7 + LayerNorm, one warp per token (register ring over rows, shuffles),
//                         written as the bf16 A operand straight into the 128B-swizzled UMMA layout in smem
//   phase 2  tcgen05      D1[128 x 2C] = A1 . W1^T, issued as NH halves of 128 hidden units
//                         (W1 TMA-staged while phase 1 runs)
//   phase 3/4, per half h CUDA cores: bias + GELU(tanh) out of TMEM -> bf16 A2_h (128 x 128) in smem (re-using
//                         A1's bytes); tcgen05: D2[128 x C] += A2_h . (gamma*W2)[:, h]^T.  GELU of half h+1
//                         overlaps the MMAs of half h.  gamma*W2 is TMA-staged into W1's bytes after phase 2.
//                         For C = 128, D2 re-uses the TMEM columns of D1's first half once GELU has drained it.
//   phase 5  CUDA cores   + gamma*b2, staged through smem (re-using A/W bytes), then coalesced out = stage + x
//
// HBM/L2 traffic per tile: x in (fp32, twice + halo rows, the second read is an L2 hit), out (fp32, once),
// weights (bf16, 4 C^2 elements).  Replaces three launches and the bf16 A16/H16 round trips.
#pragma once
#include "cnn_kernels.cuh"
#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace a2m {

#ifdef A2M_FFN_TIMING
#define FB_STAMP(i) do { if (blockIdx.x == 0 && C == 128) g_ffn_timing[(i)] = clock64(); } while (0)
#else
#define FB_STAMP(i) do { } while (0)
#endif

constexpr int FB_THREADS = 256;  // 8 warps: TMEM quadrant = warp & 3, column half = warp >> 2
constexpr int FB_TOK = 128;

template <int C>
struct FusedBlockCfg {
  static constexpr int H = 2 * C;
  static constexpr int HH = 128;                    // hidden units per half
  static constexpr int NH = H / HH;                 // 1 (C = 64) or 2 (C = 128)
  static constexpr int KB1 = C / 64;                // k-blocks of MMA1
  static constexpr int A1_BYTES = FB_TOK * C * 2;
  static constexpr int A2_BYTES = FB_TOK * HH * 2;  // one half of the hidden activations: 2 k-blocks
  static constexpr int A_BYTES = (A1_BYTES > A2_BYTES) ? A1_BYTES : A2_BYTES;   // 32 KB
  static constexpr int W_BYTES = H * C * 2;         // W1 [H, C] then W2' [C, H]: 64 KB / 16 KB
  static constexpr int STAGE_STRIDE = C + 4;        // floats; conflict-free float4 rows
  static constexpr int STAGE_BYTES = FB_TOK * STAGE_STRIDE * 4;
  static constexpr int MAIN_BYTES = (A_BYTES + W_BYTES > STAGE_BYTES) ? A_BYTES + W_BYTES : STAGE_BYTES;
  static constexpr int AUX_BYTES = (H + C) * 4 + 128;  // b1, b2', barriers, tmem slot
  static constexpr size_t SMEM = 1024 + MAIN_BYTES + AUX_BYTES;   // 99.8 KB (C = 128), 50.7 KB (C = 64)
  static constexpr uint32_t TMEM_COLS = 256;
  static constexpr uint32_t D2_COL = (C == 128) ? 0 : 128;   // C = 128: aliases D1's first half
};

// Packed fp32 parameters: dw[7][C] | dwb[C] | lnw[C] | lnb[C] | b1[H] | b2g[C] (= gamma * b2)
// TAPE (training forward): additionally writes what the backward needs -- the LN output A1 (bf16 [M, C]) and the GELU
// output A2 (bf16 [M, 2C]) by TMA store straight from their swizzled operand tiles (tmA16 / tmH16, boxes {64, 128}),
// and the pw1 pre-activation u (bf16 [M, 2C]) by per-row vector stores from the GELU phase.  The elected thread issues
// each store before the MMAs that read the same tile and waits for the store's shared-memory reads before it commits
// those MMAs to the barrier the other threads wait on, so a tile is never overwritten while a store still reads it.
template <int C, bool TAPE>
__global__ void __launch_bounds__(FB_THREADS, 2)
block_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                   const float* X, float* Y, int L, int M, const float* __restrict__ params,
                   const __grid_constant__ CUtensorMap tmA16, const __grid_constant__ CUtensorMap tmH16, __nv_bfloat16* U16) {
  using Cfg = FusedBlockCfg<C>;
  using RM = RowMap<C>;
  constexpr int H = Cfg::H;
  constexpr int PER = RM::PER;
  constexpr int NH = Cfg::NH;
  static_assert(RM::G == 1, "one vector of channels per lane");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sA = smem;                               // A1, then A2 halves
  uint8_t* sW = smem + Cfg::A_BYTES;                // W1, then W2'
  float* sStage = reinterpret_cast<float*>(smem);   // aliases sA/sW after the last MMA has completed
  float* sB1 = reinterpret_cast<float*>(smem + Cfg::MAIN_BYTES);
  float* sB2 = sB1 + H;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB2 + C);
  uint64_t* bar_w1 = bars;
  uint64_t* bar_d1 = bars + 1;
  uint64_t* bar_w2 = bars + 2;
  uint64_t* bar_m2 = bars + 3;                      // [NH]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 + NH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile0 = blockIdx.x * FB_TOK;

  pdl_launch_dependents();
  if (threadIdx.x == 0) FB_STAMP(112);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    mbar_init(bar_w1, 1);
    mbar_init(bar_d1, 1);
    mbar_init(bar_w2, 1);
    for (int i = 0; i < NH; ++i) mbar_init(&bar_m2[i], 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_w1, Cfg::W_BYTES);
#pragma unroll
    for (int kb = 0; kb < Cfg::KB1; ++kb) tma_load_2d(sW + kb * (H * 128), &tmW1, bar_w1, kb * 64, 0);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  copy_const_to_smem<(H + C) / 4, FB_THREADS>(sB1, params + 10 * C, threadIdx.x);

  // ---------------------------------------------------------------- phase 1: dwconv7 + LN -> A1
  {
    float w[7][PER], bias[PER], lw[PER], lb[PER];
#pragma unroll
    for (int t = 0; t < 7; ++t) RM::load(params + t * C, lane, w[t]);
    RM::load(params + 7 * C, lane, bias);
    RM::load(params + 8 * C, lane, lw);
    RM::load(params + 9 * C, lane, lb);
    pdl_wait();  // weights / parameters above are constants; x is produced by the previous kernel
    if (threadIdx.x == 0) FB_STAMP(113);
    constexpr int TPP = 8;                                   // tokens per pass (warp_sum8_all)
    constexpr int PASSES = FB_TOK / (FB_THREADS / 32) / TPP;  // 2
    const int col = RM::chan(lane, 0);
#pragma unroll 1
    for (int pass = 0; pass < PASSES; ++pass) {
      const int r0 = (warp * PASSES + pass) * TPP;           // first row (within the tile) of this pass
      float rows[TPP + 6][PER];                              // rows r0-3 .. r0+TPP+2 of x (this lane's channels)
#pragma unroll
      for (int i = 0; i < TPP + 6; ++i) {
        const int g = tile0 + r0 - 3 + i;
        if (g >= 0 && g < M) {
          RM::load(X + static_cast<size_t>(g) * C, lane, rows[i]);
        } else {
#pragma unroll
          for (int j = 0; j < PER; ++j) rows[i][j] = 0.f;
        }
      }
      // depthwise conv of the 8 tokens of the pass, then their LayerNorm statistics jointly (warp_sum8_all)
      float y[TPP][PER];
      const int l_first = (tile0 + r0) % L;   // one runtime modulo per pass instead of one per token (~22 instructions each)
      if (l_first >= 3 && l_first + TPP + 3 <= L) {
        // the pass and its halo lie inside one window (all but 2 of a window's 62 passes): 28 fused multiply-adds per token and
        // lane, no per-tap boundary test (the tests were a fifth of this phase's instructions; the kernel is issue-bound)
#pragma unroll
        for (int i = 0; i < TPP; ++i) {
#pragma unroll
          for (int j = 0; j < PER; ++j) y[i][j] = bias[j];
#pragma unroll
          for (int t = 0; t < 7; ++t)
#pragma unroll
            for (int j = 0; j < PER; ++j) y[i][j] = fmaf(w[t][j], rows[i + t][j], y[i][j]);
        }
      } else {
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
      }
      float st[TPP];
#pragma unroll
      for (int i = 0; i < TPP; ++i) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) a += y[i][j];
        st[i] = a;
      }
      warp_sum8_all(st, lane);
      float mean[TPP];
#pragma unroll
      for (int i = 0; i < TPP; ++i) {
        mean[i] = st[i] * (1.0f / C);
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) a += (y[i][j] - mean[i]) * (y[i][j] - mean[i]);
        st[i] = a;
      }
      warp_sum8_all(st, lane);
#pragma unroll
      for (int i = 0; i < TPP; ++i) {
        const int r = r0 + i;
        const float inv = rsqrtf(st[i] * (1.0f / C) + kLnEps);
        const bool live = tile0 + r < M;
#pragma unroll
        for (int j = 0; j < PER; ++j) y[i][j] = live ? (y[i][j] - mean[i]) * inv * lw[j] + lb[j] : 0.f;
        uint8_t* dst = sA + (col >> 6) * (FB_TOK * 128) + sw128_offset(r, col & 63);
        if constexpr (PER == 4) {
          uint2 q;
          q.x = pack_bf16x2(y[i][0], y[i][1]);
          q.y = pack_bf16x2(y[i][2], y[i][3]);
          *reinterpret_cast<uint2*>(dst) = q;
        } else {
          *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(y[i][0], y[i][1]);
        }
      }
    }
  }
  if (threadIdx.x == 0) FB_STAMP(114);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d2 = tmem_base + Cfg::D2_COL;
  if (threadIdx.x == 0) FB_STAMP(115);

  // ---------------------------------------------------------------- phase 2: D1 = A1 . W1^T (NH halves of N = 128)
  if (threadIdx.x == 0) {
    if constexpr (TAPE) {
#pragma unroll
      for (int kb = 0; kb < Cfg::KB1; ++kb) tma_store_2d(&tmA16, sA + kb * (FB_TOK * 128), kb * 64, tile0);
      bulk_commit();
    }
    mbar_wait(bar_w1, 0);
    tc_fence_after();
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, Cfg::HH);
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
#pragma unroll
      for (int kb = 0; kb < Cfg::KB1; ++kb) {
        const uint64_t da = umma_desc_sw128(smem_u32(sA + kb * (FB_TOK * 128)));
        const uint64_t db = umma_desc_sw128(smem_u32(sW + kb * (H * 128) + hh * (Cfg::HH * 128)));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + hh * Cfg::HH, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(db, k * 32), idesc1,
                    (kb | k) != 0 ? 1u : 0u);
      }
    }
    if constexpr (TAPE) bulk_wait_read<0>();
    umma_commit(bar_d1);
  }
  __syncwarp();
  mbar_wait(bar_d1, 0);
  tc_fence_after();
  if (threadIdx.x == 0) FB_STAMP(116);
  if (threadIdx.x == 0) {
    // the MMAs above have finished reading A1 and W1: stage gamma-scaled W2 into W1's bytes
    mbar_arrive_expect_tx(bar_w2, Cfg::W_BYTES);
#pragma unroll
    for (int kb = 0; kb < H / 64; ++kb) tma_load_2d(sW + kb * (C * 128), &tmW2, bar_w2, kb * 64, 0);
  }
  __syncwarp();

  // ---------------------------------------------------------------- phases 3/4 per hidden half
  const int quad = warp & 3;        // TMEM lanes 32*quad .. +31
  const int cg = warp >> 2;         // column half: 64 of the 128 hidden units of a half == one k-block of A2
  const int row = quad * 32 + lane;
  const uint32_t t_row = static_cast<uint32_t>(quad * 32) << 16;
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    uint32_t packed[32];  // 64 hidden units of this row, bf16
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col0 = hh * Cfg::HH + cg * 64 + c * 32;   // hidden unit index
      uint32_t r[32];
      tmem_ld_x32(tmem_base + t_row + col0, r);
      tmem_ld_wait();
      if constexpr (TAPE) {
        uint32_t pre[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          pre[j] = pack_bf16x2(__uint_as_float(r[2 * j]) + sB1[col0 + 2 * j], __uint_as_float(r[2 * j + 1]) + sB1[col0 + 2 * j + 1]);
        if (tile0 + row < M) {
          uint4* dst = reinterpret_cast<uint4*>(U16 + static_cast<size_t>(tile0 + row) * H + col0);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pre[4 * q], pre[4 * q + 1], pre[4 * q + 2], pre[4 * q + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        packed[c * 16 + j] = pack_bf16x2(gelu_tanh_fast(__uint_as_float(r[2 * j]) + sB1[col0 + 2 * j]),
                                         gelu_tanh_fast(__uint_as_float(r[2 * j + 1]) + sB1[col0 + 2 * j + 1]));
    }
    if (threadIdx.x == 0) FB_STAMP(117 + hh * 2);
    if (hh > 0) {
      mbar_wait(&bar_m2[hh - 1], 0);   // the previous half's MMAs have finished reading the A2 bytes
      tc_fence_after();
    }
    uint8_t* base = sA + cg * (FB_TOK * 128);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<uint4*>(base + sw128_offset(row, 8 * q)) =
          make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      if constexpr (TAPE) {
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_store_2d(&tmH16, sA + j * (FB_TOK * 128), hh * Cfg::HH + j * 64, tile0);
        bulk_commit();
      }
      if (hh == 0) mbar_wait(bar_w2, 0);
      tc_fence_after();
      FB_STAMP(118 + hh * 2);
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, C);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint64_t da = umma_desc_sw128(smem_u32(sA + j * (FB_TOK * 128)));
        const uint64_t db = umma_desc_sw128(smem_u32(sW + (2 * hh + j) * (C * 128)));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_d2, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(db, k * 32), idesc2,
                    (hh | j | k) != 0 ? 1u : 0u);
      }
      if constexpr (TAPE) bulk_wait_read<0>();
      umma_commit(&bar_m2[hh]);
    }
    __syncwarp();
  }
  // the residual rows of the first half of the store phase are requested BEFORE the wait for the last MMAs: their L2 latency
  // (9 % of the kernel's stall samples sat on the add that consumes them) hides behind the tensor core
  constexpr int NW5 = FB_THREADS / 32;
  constexpr int RPW5 = FB_TOK / NW5;  // 16 rows per warp, in two batches of 8
  constexpr int NPRE = RPW5 / 2;   // (prefetching all 16 rows was measured: no faster, and the registers spill)
  float xpre[NPRE][PER];
#pragma unroll
  for (int i = 0; i < NPRE; ++i) {
    const int tok = tile0 + warp + i * NW5;
    if (tok < M) RM::load(X + static_cast<size_t>(tok) * C, lane, xpre[i]);
  }
  mbar_wait(&bar_m2[NH - 1], 0);
  tc_fence_after();
  if (threadIdx.x == 0) FB_STAMP(121);

  // ---------------------------------------------------------------- phase 5: stage (D2 + b2'), then out = stage + x
  {
    constexpr int COLS = C / 2;  // per column half: 64 (C = 128) or 32 (C = 64)
#pragma unroll
    for (int c = 0; c < COLS / 32; ++c) {
      const int col0 = cg * COLS + c * 32;
      float* srow = sStage + row * Cfg::STAGE_STRIDE + col0;
      uint32_t r[32];
      tmem_ld_x32(tmem_d2 + t_row + col0, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        reinterpret_cast<float4*>(srow)[q] =
            make_float4(__uint_as_float(r[4 * q]) + sB2[col0 + 4 * q], __uint_as_float(r[4 * q + 1]) + sB2[col0 + 4 * q + 1],
                        __uint_as_float(r[4 * q + 2]) + sB2[col0 + 4 * q + 2], __uint_as_float(r[4 * q + 3]) + sB2[col0 + 4 * q + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) FB_STAMP(122);
  {
    // coalesced residual add: each warp owns rows warp, warp+8, ...; lanes span the channels
    constexpr int NW = FB_THREADS / 32;
    constexpr int RPW = FB_TOK / NW;  // 16 rows per warp, in two batches of 8 loads in flight
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float xv[RPW / 2][PER];
#pragma unroll
      for (int i = 0; i < RPW / 2; ++i) {
        const int tok = tile0 + warp + (half * (RPW / 2) + i) * NW;
        if (half * (RPW / 2) + i < NPRE) {
#pragma unroll
          for (int j = 0; j < PER; ++j) xv[i][j] = xpre[half * (RPW / 2) + i][j];
        } else if (tok < M) {
          RM::load(X + static_cast<size_t>(tok) * C, lane, xv[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < RPW / 2; ++i) {
        const int r = warp + (half * (RPW / 2) + i) * NW;
        const int tok = tile0 + r;
        if (tok < M) {
          float sv[PER];
          RM::load(sStage + r * Cfg::STAGE_STRIDE, lane, sv);
#pragma unroll
          for (int j = 0; j < PER; ++j) sv[j] += xv[i][j];
          RM::store_f32(Y + static_cast<size_t>(tok) * C, lane, sv);
        }
      }
    }
  }
  if (threadIdx.x == 0) FB_STAMP(123);
  if constexpr (TAPE)
    if (threadIdx.x == 0) bulk_wait_all<0>();   // the tape stores must have landed before the grid may be considered complete
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace a2m
