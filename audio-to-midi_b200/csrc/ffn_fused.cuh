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

// (round 1's ffn_fused_kernel -- LayerNorm + FFN only, the output projection as a separate GEMM -- lived here; postattn_fused.cuh
// superseded it and it was removed in round 2.  What remains is what the fused kernels share: tile constants and the LayerNorm prologue.)

}  // namespace a2m
