// Everything of a TransformerLayer after the attention core (model.py:541-553) per launch, on one tile of 128 tokens:
//     x' = x + O Wo^T                                   (SelfAttention.output_proj, residual of the attention block)
//     x  <- x' + W2 ( gelu(u[:512]) * u[512:] ) + b2,   u = W1 LN(x') + b1         (FeedForwardBlock, its residual)
// Replaces gemm_tc2<128, F32, RESID> (output projection) + ffn_fused_kernel.  The residual stream is carried through the
// launch in TENSOR MEMORY: the out-projection accumulates into D2, the compute warps add x (TMA-staged, thread-per-row)
// and write x' back into D2 with tcgen05.st, feed_forward_norm is taken from the same registers, and the FFN's second
// product accumulates on top of x' -- so x is read once and written once per launch, and the normalised / gated
// activations never leave the SM.
//
//   warp 0      TMA producer: Wo (4 stages), x (4 stages, one per 64-column quarter), then W1 / W2 as in ffn_fused.cuh,
//               all through one ring of 32 KB stages; the attention output tile O goes straight into the A-operand bytes
//   warp 1      tcgen05.mma issuer: D2 = O Wo^T; then per hidden chunk MMA1 -> D1[c & 1], MMA2 accumulating into D2
//   warps 2-17  (a) x' = D2 + x, LayerNorm statistics across the 4 column-quarter warps of a row, x' -> TMEM, LN(x') -> A
//               (b) gate every hidden chunk (bias, gelu, product) out of TMEM -> bf16 h chunk
//               (c) D2 + b2 staged through shared memory, coalesced store of x
#pragma once
#include "ffn_fused.cuh"

namespace a2m {

#ifndef A2M_PA_TMA_STORE
#define A2M_PA_TMA_STORE 1
#endif

constexpr int PA_AUX_BYTES = (2 * FF_F + FF_D) * 4 + 2 * FF_D * 4 + 2 * FF_ROWS * 4 * 4 + 256;   // b1, b2 | lnw, lnb | row stats | barriers
constexpr size_t PA_SMEM = 1024 + FF_MAIN_BYTES + PA_AUX_BYTES;

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tmO: attention output [M, 256] bf16, box {64, 128};  tmWo: Wo [256, 256] bf16, box {64, 256};
// tmX: x [M, 256] fp32, box {32, 128} (128B swizzle);  tmW1 / tmW2 / b1p / b2 / lnw / lnb as in ffn_fused_kernel.
__global__ void __launch_bounds__(FF_THREADS, 1)
postattn_fused_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWo,
                      const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                      const __grid_constant__ CUtensorMap tmW2, float* X, int M, const float* __restrict__ lnw,
                      const float* __restrict__ lnb, const float* __restrict__ b1p, const float* __restrict__ b2) {
  using RM = RowMap<FF_D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sH = sA + FF_A_BYTES;
  uint8_t* sW = sH + 2 * FF_H_BYTES;
#if !A2M_PA_TMA_STORE
  float* sStage = reinterpret_cast<float*>(smem);   // aliases everything above once the last MMA has completed
#endif
  float* sB1 = reinterpret_cast<float*>(smem + FF_MAIN_BYTES);
  float* sB2 = sB1 + 2 * FF_F;
  float* sLnW = sB2 + FF_D;
  float* sLnB = sLnW + FF_D;
  float* sStat = sLnB + FF_D;               // [2][128 rows][4 quarters]
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sStat + 2 * FF_ROWS * 4);
  uint64_t* bar_empty = bar_full + FF_NST;
  uint64_t* bar_a = bar_empty + FF_NST;     // A operand (LayerNorm output) ready, x' in TMEM
  uint64_t* bar_d1full = bar_a + 1;         // [2]
  uint64_t* bar_d1free = bar_d1full + 2;    // [2]
  uint64_t* bar_hfull = bar_d1free + 2;     // [2]
  uint64_t* bar_hfree = bar_hfull + 2;      // [2]
  uint64_t* bar_done = bar_hfree + 2;
  uint64_t* bar_o = bar_done + 1;           // attention output tile landed in sA
  uint64_t* bar_p = bar_o + 1;              // out-projection accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_p + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile0 = blockIdx.x * FF_ROWS;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmWo);
    tma_prefetch_desc(&tmX);
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
    mbar_init(bar_o, 1);
    mbar_init(bar_p, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (threadIdx.x < (2 * FF_F + FF_D) / 4) {   // 320 float4: b1 (packed) then b2
    const int i = threadIdx.x;
    reinterpret_cast<float4*>(sB1)[i] = (i < 2 * FF_F / 4) ? __ldg(reinterpret_cast<const float4*>(b1p) + i)
                                                           : __ldg(reinterpret_cast<const float4*>(b2) + i - 2 * FF_F / 4);
  } else if (threadIdx.x < (2 * FF_F + FF_D) / 4 + 2 * FF_D / 4) {   // 128 float4: lnw then lnb
    const int i = threadIdx.x - (2 * FF_F + FF_D) / 4;
    reinterpret_cast<float4*>(sLnW)[i] = (i < FF_D / 4) ? __ldg(reinterpret_cast<const float4*>(lnw) + i)
                                                        : __ldg(reinterpret_cast<const float4*>(lnb) + i - FF_D / 4);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d2 = tmem_base;             // columns 0..255: out-projection, then x', then x' + FFN
  const uint32_t tmem_d1 = tmem_base + 256;       // two accumulators of 128 columns

  if (warp == 0) {
    // ------------------------------------------------------------ producer
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      auto next = [&]() { if (++s == FF_NST) { s = 0; ph ^= 1; } };
      auto load_wo = [&](int kb) {
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], FF_STAGE);
        tma_load_2d(sW + s * FF_STAGE, &tmWo, &bar_full[s], kb * 64, 0);
        next();
      };
      for (int kb = 0; kb < FF_NST; ++kb) load_wo(kb);   // constants: issued before the dependency wait
      pdl_wait();                                        // O and x are produced by the previous kernels
      mbar_arrive_expect_tx(bar_o, FF_A_BYTES);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(sA + kb * (FF_ROWS * 128), &tmO, bar_o, kb * 64, tile0);
      for (int kb = FF_NST; kb < 4; ++kb) load_wo(kb);
      for (int q = 0; q < 4; ++q) {                      // x, one stage per 64-column quarter (two boxes of 32 fp32 columns)
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], FF_STAGE);
        tma_load_2d(sW + s * FF_STAGE, &tmX, &bar_full[s], q * 64, tile0);
        tma_load_2d(sW + s * FF_STAGE + 16384, &tmX, &bar_full[s], q * 64 + 32, tile0);
        next();
      }
      auto load_w1 = [&](int c, int half) {
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], FF_STAGE);
        tma_load_2d(sW + s * FF_STAGE, &tmW1, &bar_full[s], (2 * half) * 64, c * 128);
        tma_load_2d(sW + s * FF_STAGE + 16384, &tmW1, &bar_full[s], (2 * half + 1) * 64, c * 128);
        next();
      };
      auto load_w2 = [&](int c) {
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], FF_STAGE);
        tma_load_2d(sW + s * FF_STAGE, &tmW2, &bar_full[s], c * FF_CH, 0);
        next();
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
      auto next = [&]() { if (++s == FF_NST) { s = 0; ph ^= 1; } };
      // D2 = O Wo^T
      mbar_wait(bar_o, 0);
      for (int kb = 0; kb < 4; ++kb) {
        mbar_wait(&bar_full[s], ph);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem_u32(sA + kb * (FF_ROWS * 128)));
        const uint64_t db = umma_desc_sw128(smem_u32(sW + s * FF_STAGE));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_d2, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(db, k * 32), idesc2, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&bar_empty[s]);
        next();
      }
      umma_commit(bar_p);
      for (int q = 0; q < 4; ++q) next();   // the four x stages are consumed by the compute warps
      mbar_wait(bar_a, 0);
      tc_fence_after();
      for (int c = 0; c <= FF_NCH; ++c) {
        if (c < FF_NCH) {
          mbar_wait(&bar_d1free[c & 1], ((c >> 1) & 1) ^ 1);
          tc_fence_after();
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
            next();
          }
          umma_commit(&bar_d1full[c & 1]);
        }
        if (c >= 1) {
          const int cc = c - 1;
          mbar_wait(&bar_hfull[cc & 1], (cc >> 1) & 1);
          mbar_wait(&bar_full[s], ph);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(sH + (cc & 1) * FF_H_BYTES));
          const uint64_t db = umma_desc_sw128(smem_u32(sW + s * FF_STAGE));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // always accumulating: D2 already holds x'
            umma_bf16(tmem_d2, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(db, k * 32), idesc2, 1u);
          umma_commit(&bar_empty[s]);
          umma_commit(&bar_hfree[cc & 1]);
          next();
        }
      }
      umma_commit(bar_done);
    }
  } else {
    // ------------------------------------------------------------ compute warps
    const int cw = warp - 2;            // 0..15
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int cq = cw >> 2;             // column quarter
    const int row = quad * 32 + lane;
    const uint32_t t_row = static_cast<uint32_t>(quad * 32) << 16;

    // (a) x' = O Wo^T + x ; statistics ; x' -> TMEM ; LN(x') -> A operand k-block cq
    {
      const int item = 4 + cq;                        // ring position of this quarter's x stage
      const int xs = item % FF_NST, xph = (item / FF_NST) & 1;
      float v[64];
      mbar_wait(bar_p, 0);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld_x32(tmem_d2 + t_row + cq * 64 + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[half * 32 + j] = __uint_as_float(r[j]);
      }
      // a waiter must observe every phase of a barrier in order: stage xs was filled (item - 3) by a Wo block, which has
      // completed (bar_p), but quarter 3's stage also carried quarter 0's x (item 4) in between
      if (cq == 3) mbar_wait(&bar_full[xs], xph ^ 1);
      mbar_wait(&bar_full[xs], xph);
      float sum = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint8_t* xrow = sW + xs * FF_STAGE + half * 16384 + row * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 x4 = *reinterpret_cast<const float4*>(xrow + ((static_cast<uint32_t>(q) ^ (row & 7)) << 4));
          float* d = v + half * 32 + 4 * q;
          d[0] += x4.x; d[1] += x4.y; d[2] += x4.z; d[3] += x4.w;
          sum += (d[0] + d[1]) + (d[2] + d[3]);
        }
      }
      // the stage may be refilled once all 128 threads of this quarter have read it
      named_bar_sync(4 + cq, 128);
      if ((cw & 3) == 0 && lane == 0) mbar_arrive(&bar_empty[xs]);
      // x' back into D2 (the FFN's second product accumulates onto it)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(v[half * 32 + j]);
        tmem_st_x32(tmem_d2 + t_row + cq * 64 + half * 32, r);
      }
      sStat[row * 4 + cq] = sum;
      named_bar_sync(1, FF_CTHREADS);
      const float4 s4 = *reinterpret_cast<const float4*>(sStat + row * 4);
      const float mean = ((s4.x + s4.y) + (s4.z + s4.w)) * (1.0f / FF_D);
      float var = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) var += (v[j] - mean) * (v[j] - mean);
      sStat[FF_ROWS * 4 + row * 4 + cq] = var;
      named_bar_sync(2, FF_CTHREADS);
      const float4 q4 = *reinterpret_cast<const float4*>(sStat + FF_ROWS * 4 + row * 4);
      const float inv = rsqrtf(((q4.x + q4.y) + (q4.z + q4.w)) * (1.0f / FF_D) + kLnEps);
      uint8_t* ablk = sA + cq * (FF_ROWS * 128);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = cq * 64 + 8 * q + j;
          y[j] = (v[8 * q + j] - mean) * inv * sLnW[col] + sLnB[col];
        }
        *reinterpret_cast<uint4*>(ablk + sw128_offset(row, 8 * q)) =
            make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_a);
    }

    // (b) gate every hidden chunk: this thread owns 16 of the chunk's 64 hidden units of its row
#pragma unroll 1
    for (int c = 0; c < FF_NCH; ++c) {
      mbar_wait(&bar_d1full[c & 1], (c >> 1) & 1);
      tc_fence_after();
      const uint32_t d1 = tmem_d1 + (c & 1) * 128 + t_row;
      uint32_t r1[16], r2[16];
      tmem_ld_x16(d1 + cq * 16, r1);
      tmem_ld_x16(d1 + 64 + cq * 16, r2);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bar_d1free[c & 1]);
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
    }

    // (c) D2 (= x' + FFN) + b2 -> staging (all operand bytes are dead once bar_done fires), then coalesced store
    mbar_wait(bar_done, 0);
    tc_fence_after();
#if A2M_PA_TMA_STORE
    // eight [128 rows x 32 fp32] boxes in tmX's 128B-swizzled layout (the layout the x stages arrived in), handed back to global
    // memory by eight TMA stores: 256 LDS + 256 STG warp instructions per CTA less (the store loop was throttled by the LSU)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col0 = cq * 64 + c * 32;
      uint32_t r[32];
      tmem_ld_x32(tmem_d2 + t_row + col0, r);
      tmem_ld_wait();
      uint8_t* brow = smem + (col0 >> 5) * 16384 + row * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(brow + ((static_cast<uint32_t>(q) ^ (row & 7)) << 4)) =
            make_float4(__uint_as_float(r[4 * q]) + sB2[col0 + 4 * q], __uint_as_float(r[4 * q + 1]) + sB2[col0 + 4 * q + 1],
                        __uint_as_float(r[4 * q + 2]) + sB2[col0 + 4 * q + 2], __uint_as_float(r[4 * q + 3]) + sB2[col0 + 4 * q + 3]);
    }
    fence_proxy_async_smem();
    named_bar_sync(1, FF_CTHREADS);
    if (threadIdx.x == 64) {
#pragma unroll
      for (int bx = 0; bx < 8; ++bx) tma_store_2d(&tmX, smem + bx * 16384, bx * 32, tile0);
      bulk_commit();
      bulk_wait_all<0>();   // the stores must have landed before the grid counts as complete
    }
  }

#else
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
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = cw + i * FF_CWARPS;
      if (tile0 + r < M) {
        float sv[RM::PER];
        RM::load(sStage + r * FF_STAGE_STRIDE, lane, sv);
        RM::store_f32(X + static_cast<size_t>(tile0 + r) * FF_D, lane, sv);
      }
    }
  }

#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace a2m
