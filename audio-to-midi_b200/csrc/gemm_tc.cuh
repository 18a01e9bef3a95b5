// Persistent, warp-specialised tcgen05 GEMM for the model's dense contractions:
//     D[M, N] = A[M, K] (bf16, K-major)  x  W[N, K]^T (bf16, K-major)   -> fp32 in TMEM
// with the memory-bound glue of the reference layers fused into the epilogue.
//
// Serves (reference call sites): Block.point_conv_1/2 (model.py:163-166), Downsample.conv
// (model.py:118), SelfAttention q/kv/k/v/out projections (model.py:353-369), FeedForwardBlock
// (model.py:232-236) and Decoder.decoder_pooling + sigmoid (model.py:192-193).
//
// Structure (one CTA = 6 warps, persistent over output tiles of 128 x BN):
//   warp 0    TMA producer: A tile 128x64 and W tile BNx64 per k-block, 128B-swizzled, 4-stage ring
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16)
//   warps 2-5 epilogue: tcgen05.ld of the fp32 accumulator (one row per thread), fused math, stores
// TMEM holds two accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the MMAs of
// tile i+1.  All waits are mbarrier waits with a trap-on-timeout (ptx.cuh).
#pragma once
#include "ptx.cuh"

namespace a2m {

enum GemmMode : int {
  GEMM_GENERIC = 0,  // [+bias] [gelu] [*gamma] [+resid] -> fp32 and/or bf16
  GEMM_GLU = 1,      // gelu(x1 + b1) * (x2 + b2) with x1/x2 interleaved per tile -> bf16
  GEMM_ROPE = 2,     // RoPE on columns < rope_cols -> bf16; columns >= vt_col0 optionally stored transposed
  GEMM_DECODER = 3,  // +bias -> logits, sigmoid -> probs; compact (B, 250, 90) fp32
};

enum GemmFlags : uint32_t {
  GF_BIAS = 1u, GF_GELU = 2u, GF_GAMMA = 4u, GF_RESID = 8u, GF_OUT32 = 16u, GF_OUT16 = 32u,
};

struct GemmArgs {
  int M, N, K;
  uint32_t flags;
  const float* bias;
  const float* gamma;
  const float* resid;
  int ldr;
  float* out32;
  int ld32;
  __nv_bfloat16* out16;
  int ld16;
  // GEMM_ROPE
  const float* rope_cos;
  const float* rope_sin;
  int rope_cols;
  int rows_per_window;  // position = row % rows_per_window
  __nv_bfloat16* vt_out;  // [B, heads, 64, rows_per_window] when non-null
  int vt_col0;
  // GEMM_DECODER
  float* logits;      // any of the three may be null
  float* probs;
  __half* probs16;    // probabilities as IEEE binary16 (the host path's compact output, a2m_submit_host_ex)
  int valid_rows;  // 250
  int valid_cols;  // 90
  // training: dropout of (acc + bias) before the residual add (FeedForwardBlock, model.py:237); gemm_tc2 G2_F32 only
  const DropParams* drop;
  uint32_t drop_site;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 192;

template <int BN>
constexpr size_t gemm_smem_bytes() {
  return 1024 + GEMM_STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + 256;
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh(u)) == x * sigmoid(2u), u = sqrt(2/pi) (x + 0.044715 x^3)   [jax.nn.gelu default]
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return __fdividef(x, 1.0f + __expf(-2.0f * u));
}
// Same function through the hardware tanh (one MUFU op instead of ex2 + rcp; |error| <= 2^-11 relative on tanh, below the
// bf16 rounding of every consumer).  Used where the activation is the per-chunk critical path (ffn_fused.cuh).
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float x2 = x * x;
  const float u = x * fmaf(0.0356774081f, x2, 0.7978845608f);   // sqrt(2/pi) (x + 0.044715 x^3)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = op2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Stores 32 consecutive fp32 / bf16 values of one row (16-byte aligned destinations).
__device__ __forceinline__ void store_row32_f32(float* dst, const float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* dst, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 q;
    q.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
    q.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
    q.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
    q.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
    reinterpret_cast<uint4*>(dst)[j] = q;
  }
}

template <int BN, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = BN * GEMM_BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sA = smem;
  uint8_t* sB = smem + GEMM_STAGES * A_BYTES;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB + GEMM_STAGES * B_BYTES);
  uint64_t* bar_empty = bar_full + GEMM_STAGES;
  uint64_t* bar_tfull = bar_empty + GEMM_STAGES;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_n = g.N / BN;
  const int num_m = (g.M + GEMM_BM - 1) / GEMM_BM;
  const int num_tiles = num_m * num_n;
  const int num_kb = g.K / GEMM_BK;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_tfull[a], 1);
      mbar_init(&bar_tempty[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&bar_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&bar_full[s], A_BYTES + B_BYTES);
          tma_load_2d(sA + s * A_BYTES, &tmA, &bar_full[s], kb * GEMM_BK, m_blk * GEMM_BM);
          tma_load_2d(sB + s * B_BYTES, &tmB, &bar_full[s], kb * GEMM_BK, n_blk * BN);
          if (++s == GEMM_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      uint32_t s = 0, ph = 0, a = 0, aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&bar_tempty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&bar_full[s], ph);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(sA + s * A_BYTES));
          const uint64_t db = umma_desc_sw128(smem_u32(sB + s * B_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16(d_tmem, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(db, k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(&bar_empty[s]);  // frees the smem slot once these MMAs have read it
          if (++s == GEMM_STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&bar_tfull[a]);  // accumulator complete
        if (++a == 2) { a = 0; aph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    uint32_t a = 0, aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      mbar_wait(&bar_tfull[a], aph);
      tc_fence_after();
      const int row = m_blk * GEMM_BM + quad * 32 + lane;
      const bool row_ok = row < g.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + a * BN;

      if constexpr (MODE == GEMM_GENERIC) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int col0 = n_blk * BN + c * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (g.flags & GF_BIAS) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col0) + j);
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if (g.flags & GF_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
          }
          if (g.flags & GF_GAMMA) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(g.gamma + col0) + j);
              v[4 * j] *= b.x; v[4 * j + 1] *= b.y; v[4 * j + 2] *= b.z; v[4 * j + 3] *= b.w;
            }
          }
          if (row_ok) {
            if (g.flags & GF_RESID) {
              const float4* rp = reinterpret_cast<const float4*>(g.resid + static_cast<size_t>(row) * g.ldr + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = rp[j];
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
              }
            }
            if (g.flags & GF_OUT32) store_row32_f32(g.out32 + static_cast<size_t>(row) * g.ld32 + col0, v);
            if (g.flags & GF_OUT16) store_row32_bf16(g.out16 + static_cast<size_t>(row) * g.ld16 + col0, v);
          }
        }
      } else if constexpr (MODE == GEMM_GLU) {
        constexpr int HALF = BN / 2;
#pragma unroll 1
        for (int c = 0; c < HALF / 32; ++c) {
          uint32_t r1[32], r2[32];
          tmem_ld_x32(taddr + c * 32, r1);
          tmem_ld_x32(taddr + HALF + c * 32, r2);
          tmem_ld_wait();
          const float* b1 = g.bias + n_blk * BN + c * 32;
          const float* b2 = b1 + HALF;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x1 = __uint_as_float(r1[j]) + __ldg(b1 + j);
            const float x2 = __uint_as_float(r2[j]) + __ldg(b2 + j);
            v[j] = gelu_tanh_f(x1) * x2;
          }
          if (row_ok) store_row32_bf16(g.out16 + static_cast<size_t>(row) * g.ld16 + n_blk * HALF + c * 32, v);
        }
      } else if constexpr (MODE == GEMM_ROPE) {
        const int pos = row % g.rows_per_window;
        const int win = row / g.rows_per_window;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int col0 = n_blk * BN + c * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (col0 < g.rope_cols) {  // rope.py:43-52: pairs (2i, 2i+1), i = (col % 64) / 2
            const int i0 = (col0 & 63) >> 1;
            const float4* cp = reinterpret_cast<const float4*>(g.rope_cos + pos * 32 + i0);
            const float4* sp = reinterpret_cast<const float4*>(g.rope_sin + pos * 32 + i0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 cs = __ldg(cp + j), sn = __ldg(sp + j);
              const float cc[4] = {cs.x, cs.y, cs.z, cs.w};
              const float ss[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float x1 = v[8 * j + 2 * t], x2 = v[8 * j + 2 * t + 1];
                v[8 * j + 2 * t] = x1 * cc[t] - x2 * ss[t];
                v[8 * j + 2 * t + 1] = x1 * ss[t] + x2 * cc[t];
              }
            }
          }
          if (row_ok) {
            if (g.vt_out != nullptr && col0 >= g.vt_col0) {
              // V stored transposed per (window, head): [win][h][d][pos], so that P.V reads V^T K-major
              const int hc = col0 - g.vt_col0;  // = h * 64 + d0
              __nv_bfloat16* dst = g.vt_out + (static_cast<size_t>(win) * 256 + hc) * g.rows_per_window + pos;
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[static_cast<size_t>(j) * g.rows_per_window] = op1_rn(v[j]);
            } else {
              store_row32_bf16(g.out16 + static_cast<size_t>(row) * g.ld16 + col0, v);
            }
          }
        }
      } else {  // GEMM_DECODER
        const int pos = row % g.rows_per_window;
        const int win = row / g.rows_per_window;
        const bool ok = row_ok && pos < g.valid_rows;
        const size_t obase = (static_cast<size_t>(win) * g.valid_rows + pos) * g.valid_cols;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int col0 = n_blk * BN + c * 32;
          if (ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              if (col < g.valid_cols) {
                const float z = __uint_as_float(r[j]) + __ldg(g.bias + col);
                const float pr = sigmoid_f(z);
                if (g.logits) g.logits[obase + col] = z;
                if (g.probs) g.probs[obase + col] = pr;
                if (g.probs16) g.probs16[obase + col] = __float2half_rn(pr);
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bar_tempty[a]);
      if (++a == 2) { a = 0; aph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace a2m
