// gemm_tc2_kernel: the production tcgen05 GEMM.  Same mainloop as gemm_tc_kernel (gemm_tc.cuh: TMA ring ->
// tcgen05.mma -> double-buffered TMEM accumulator, persistent over 128 x BN tiles) with an epilogue that
// never touches global memory from the epilogue threads:
//
//   * bias / layer-scale vectors are copied to shared memory once per CTA;
//   * the fp32 residual tile (tensor map tmR; may be the output itself for an in-place update, or another buffer
//     when the training forward must keep its input) is TMA-loaded by a dedicated producer warp into a ring of 16 KB staging chunks
//     (128 rows x 128 B, 128B-swizzled) while the MMAs of the tile run;
//   * each epilogue thread (one accumulator row) adds its row chunk in shared memory, in place;
//   * one elected thread TMA-stores the chunk (bulk async group); rows beyond M are clipped by the tensor map.
//
// The first version (gemm_tc.cuh) issued per-thread, row-strided global loads/stores and three dependent
// global round trips per 32-column chunk; ncu showed >80 % of its stall samples on those (profiles/r01a_*).
#pragma once
#include "gemm_tc.cuh"

namespace a2m {

// d/dx gelu_tanh(x) (the arithmetic of train_kernels.cuh gelu_tanh_grad, derivative only)
__device__ __forceinline__ float gelu_tanh_deriv(float x) {
  const float k = 0.7978845608028654f, a = 0.044715f;
  const float u = k * (x + a * x * x * x);
  const float s = __fdividef(1.0f, 1.0f + __expf(-2.0f * u));
  return s + x * s * (1.0f - s) * 2.0f * k * (1.0f + 3.0f * a * x * x);
}

enum Gemm2Mode : int {
  G2_F32 = 0,   // [+bias] [gelu] [*gamma] [+resid (template)] -> fp32
  G2_BF16 = 1,  // [+bias] [gelu] -> bf16
  G2_GLU = 2,   // gelu(x1 + b1) * (x2 + b2) -> bf16, N/2 output columns
  G2_ROPE = 3,  // RoPE on columns < rope_cols -> bf16; columns >= vt_col0 stored transposed (direct)
};

constexpr int G2_THREADS = 224;  // warps: 0 TMA operands, 1 MMA, 2 TMA residual, 3-6 epilogue
constexpr int G2_NCH = 4;        // staging chunks in flight
constexpr int G2_CHUNK = 128 * 128;
constexpr int G2_MAXN = 1024;

template <int BN>
__host__ __device__ constexpr int g2_stages() { return BN == 256 ? 3 : 4; }
template <int BN>
constexpr size_t gemm2_smem_bytes() {
  return 1024 + g2_stages<BN>() * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + G2_NCH * G2_CHUNK + 2 * G2_MAXN * 4 + 256;
}

template <int BN, int MODE, bool RESID>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmArgs g) {
  // RESID: G2_F32 -- fp32 residual tile added in place
  static_assert(!RESID || MODE == G2_F32, "the staged second operand exists for the fp32 epilogue");
  constexpr int STAGES = g2_stages<BN>();
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = BN * GEMM_BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * A_BYTES;
  uint8_t* sStage = sB + STAGES * B_BYTES;
  float* sBias = reinterpret_cast<float*>(sStage + G2_NCH * G2_CHUNK);
  float* sGamma = sBias + G2_MAXN;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sGamma + G2_MAXN);
  uint64_t* bar_empty = bar_full + STAGES;
  uint64_t* bar_tfull = bar_empty + STAGES;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint64_t* bar_cfull = bar_tempty + 2;
  uint64_t* bar_cempty = bar_cfull + G2_NCH;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_cempty + G2_NCH);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_n = g.N / BN;
  const int num_m = (g.M + GEMM_BM - 1) / GEMM_BM;
  const int num_tiles = num_m * num_n;
  const int num_kb = g.K / GEMM_BK;
  // Tile schedule (identical in every role).  RoPE mode: CTAs with even / odd blockIdx own the even / odd
  // m-blocks, so the position of a thread's row (row % 256) -- and with it the cos/sin row it keeps in
  // registers -- never changes.  The host launches an even grid and num_m is even (two m-blocks per window).
  auto tile_coords = [&](int it, int& m_blk, int& n_blk) -> bool {
    if constexpr (MODE == G2_ROPE) {
      const int li = static_cast<int>(blockIdx.x >> 1) + it * static_cast<int>(gridDim.x >> 1);
      if (li >= (num_m >> 1) * num_n) return false;
      m_blk = 2 * (li / num_n) + static_cast<int>(blockIdx.x & 1);
      n_blk = li % num_n;
      return true;
    } else {
      const int tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      if (tile >= num_tiles) return false;
      m_blk = tile / num_n;
      n_blk = tile % num_n;
      return true;
    }
  };

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if constexpr (RESID) tma_prefetch_desc(&tmR);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_tfull[a], 1);
      mbar_init(&bar_tempty[a], 128);
    }
    for (int c = 0; c < G2_NCH; ++c) {
      mbar_init(&bar_cfull[c], 1);
      mbar_init(&bar_cempty[c], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  {
    // bias / layer-scale vectors (constants, N <= 1024, N % 4 == 0): both loads of a thread are issued before its stores
    const bool want_b = (g.flags & GF_BIAS) || MODE == G2_GLU, want_g = (g.flags & GF_GAMMA) != 0;
    const int nv = g.N >> 2;
    float4 vb[2], vg[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = threadIdx.x + k * G2_THREADS;
      if (want_b && i < nv) vb[k] = __ldg(reinterpret_cast<const float4*>(g.bias) + i);
      if (want_g && i < nv) vg[k] = __ldg(reinterpret_cast<const float4*>(g.gamma) + i);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = threadIdx.x + k * G2_THREADS;
      if (want_b && i < nv) reinterpret_cast<float4*>(sBias)[i] = vb[k];
      if (want_g && i < nv) reinterpret_cast<float4*>(sGamma)[i] = vg[k];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above touched only constants (weights, bias); activations are read below

  if (warp == 0) {
    // ------------------------------------------------------------ operand producer
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      int m_blk, n_blk;
      for (int it = 0; tile_coords(it, m_blk, n_blk); ++it) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&bar_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&bar_full[s], A_BYTES + B_BYTES);
          tma_load_2d(sA + s * A_BYTES, &tmA, &bar_full[s], kb * GEMM_BK, m_blk * GEMM_BM);
          tma_load_2d(sB + s * B_BYTES, &tmB, &bar_full[s], kb * GEMM_BK, n_blk * BN);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      uint32_t s = 0, ph = 0, a = 0, aph = 0;
      int m_blk, n_blk;
      for (int it = 0; tile_coords(it, m_blk, n_blk); ++it) {
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
          umma_commit(&bar_empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&bar_tfull[a]);
        if (++a == 2) { a = 0; aph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ residual producer (fp32 tile chunks)
    if constexpr (RESID) {
      if (elect_one()) {
        uint32_t ck = 0;
        int m_blk, n_blk;
        for (int it = 0; tile_coords(it, m_blk, n_blk); ++it) {
          constexpr int CW = (MODE == G2_F32) ? 32 : 64;   // columns per 128-byte staging row
          for (int c = 0; c < BN / CW; ++c, ++ck) {
            const uint32_t cb = ck % G2_NCH, cph = (ck / G2_NCH) & 1;
            mbar_wait(&bar_cempty[cb], cph ^ 1);
            mbar_arrive_expect_tx(&bar_cfull[cb], G2_CHUNK);
            tma_load_2d(sStage + cb * G2_CHUNK, &tmR, &bar_cfull[cb], n_blk * BN + c * CW, m_blk * GEMM_BM);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 3..6)
    const int quad = warp & 3;
    const int epi_tid = threadIdx.x - 96;
    const int r = quad * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t rsw = static_cast<uint32_t>(r & 7);
    uint32_t a = 0, aph = 0, ck = 0;
    uint32_t drop_k = 0u, drop_thresh = 0u;
    float drop_inv = 1.f;
    if (MODE == G2_F32 && g.drop != nullptr) {
      drop_thresh = g.drop->thresh;
      drop_inv = g.drop->inv_keep;
      drop_k = drop_key(g.drop->seed, g.drop_site);
    }
    float rc[MODE == G2_ROPE ? 32 : 1], rs[MODE == G2_ROPE ? 32 : 1];  // this row's cos / sin (all 32 pairs)
    int rope_mblk = -1;
    int m_blk, n_blk;
    for (int it = 0; tile_coords(it, m_blk, n_blk); ++it) {
      const int row0 = m_blk * GEMM_BM;
      if constexpr (MODE == G2_ROPE) {
        if ((m_blk & 1) != rope_mblk) {  // once per CTA (see tile_coords); before the accumulator wait
          rope_mblk = m_blk & 1;
          const int pos = row0 % g.rows_per_window + r;
          const float4* cp = reinterpret_cast<const float4*>(g.rope_cos + pos * 32);
          const float4* sp = reinterpret_cast<const float4*>(g.rope_sin + pos * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 cs = __ldg(cp + j), sn = __ldg(sp + j);
            rc[4 * j] = cs.x; rc[4 * j + 1] = cs.y; rc[4 * j + 2] = cs.z; rc[4 * j + 3] = cs.w;
            rs[4 * j] = sn.x; rs[4 * j + 1] = sn.y; rs[4 * j + 2] = sn.z; rs[4 * j + 3] = sn.w;
          }
        }
      }
      mbar_wait(&bar_tfull[a], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + a * BN;

      if constexpr (MODE == G2_F32) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c, ++ck) {
          const uint32_t cb = ck % G2_NCH;
          uint8_t* srow = sStage + cb * G2_CHUNK + r * 128;
          uint32_t acc[32];
          tmem_ld_x32(taddr + c * 32, acc);
          if constexpr (RESID) mbar_wait(&bar_cfull[cb], (ck / G2_NCH) & 1);
          tmem_ld_wait();
          const int col0 = n_blk * BN + c * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
          if (g.flags & GF_BIAS) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += sBias[col0 + j];
          }
          if (g.flags & GF_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
          }
          if (g.flags & GF_GAMMA) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= sGamma[col0 + j];
          }
          if (g.drop != nullptr && drop_thresh != 0u) {
            const uint32_t base = static_cast<uint32_t>(row0 + r) * static_cast<uint32_t>(g.N) + static_cast<uint32_t>(col0);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= drop_mul(drop_k, base + j, drop_thresh, drop_inv);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4* p = reinterpret_cast<float4*>(srow + ((static_cast<uint32_t>(q) ^ rsw) << 4));
            float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            if constexpr (RESID) {
              const float4 x = *p;
              o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
            }
            *p = o;
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (epi_tid == 0) {
            tma_store_2d(&tmC, sStage + cb * G2_CHUNK, col0, row0);
            bulk_commit();
            bulk_wait_read<1>();  // every store but the one just issued has finished reading its chunk
            if constexpr (RESID)
              if (ck > 0) mbar_arrive(&bar_cempty[(ck - 1) % G2_NCH]);
          }
        }
      } else {
        // bf16 outputs: one staging chunk = 64 output columns
        constexpr int OUT_COLS = (MODE == G2_GLU) ? BN / 2 : BN;
        const int pos = (MODE == G2_ROPE) ? row0 % g.rows_per_window + r : 0;  // tiles never straddle windows
        const int win = (MODE == G2_ROPE) ? row0 / g.rows_per_window : 0;
#pragma unroll 1
        for (int c2 = 0; c2 < OUT_COLS / 64; ++c2) {
          const int ocol0 = n_blk * OUT_COLS + c2 * 64;  // first output column of this chunk
          const bool transposed = (MODE == G2_ROPE) && g.vt_out != nullptr && ocol0 >= g.vt_col0;
          const uint32_t cb = ck % G2_NCH;
          uint8_t* srow = sStage + cb * G2_CHUNK + r * 128;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float v[32];
            if constexpr (MODE == G2_GLU) {
              uint32_t r1[32], r2[32];
              tmem_ld_x32(taddr + c2 * 64 + half * 32, r1);
              tmem_ld_x32(taddr + BN / 2 + c2 * 64 + half * 32, r2);
              tmem_ld_wait();
              const float* b1 = sBias + n_blk * BN + c2 * 64 + half * 32;
              const float* b2 = b1 + BN / 2;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                v[j] = gelu_tanh_f(__uint_as_float(r1[j]) + b1[j]) * (__uint_as_float(r2[j]) + b2[j]);
            } else {
              uint32_t acc[32];
              tmem_ld_x32(taddr + c2 * 64 + half * 32, acc);
              tmem_ld_wait();
              const int col0 = ocol0 + half * 32;
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
              if constexpr (MODE == G2_BF16) {
                if (g.flags & GF_BIAS) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] += sBias[col0 + j];
                }
                if (g.flags & GF_GELU) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
                }
              } else if (col0 < g.rope_cols) {  // rope.py:43-52, pairs (2i, 2i+1), i = (col % 64) / 2
                constexpr int kPairsPerHalf = 16;  // output chunks start at multiples of 64 columns
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                  const float x1 = v[2 * t], x2 = v[2 * t + 1];
                  const float cc = rc[half * kPairsPerHalf + t], ss = rs[half * kPairsPerHalf + t];
                  v[2 * t] = x1 * cc - x2 * ss;
                  v[2 * t + 1] = x1 * ss + x2 * cc;
                }
              }
              if (transposed) {
                // V^T per (window, head): [win][h * 64 + d][pos]; a warp writes 32 consecutive positions
                if (row0 + r < g.M) {
                  const int hc = col0 - g.vt_col0;
                  __nv_bfloat16* dst = g.vt_out + (static_cast<size_t>(win) * 256 + hc) * g.rows_per_window + pos;
#pragma unroll
                  for (int j = 0; j < 32; ++j) dst[static_cast<size_t>(j) * g.rows_per_window] = op1_rn(v[j]);
                }
                continue;
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 o;
              o.x = pack_bf16x2(v[8 * q], v[8 * q + 1]);
              o.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
              o.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
              o.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
              *reinterpret_cast<uint4*>(srow + ((static_cast<uint32_t>(half * 4 + q) ^ rsw) << 4)) = o;
            }
          }
          if (transposed) continue;  // uniform over the epilogue warps: no staging chunk was used
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (epi_tid == 0) {
            tma_store_2d(&tmC, sStage + cb * G2_CHUNK, ocol0, row0);
            bulk_commit();
            bulk_wait_read<1>();
            if constexpr (RESID)
              if (ck > 0) mbar_arrive(&bar_cempty[(ck - 1) % G2_NCH]);
          }
          ++ck;
        }
      }
      tc_fence_before();
      mbar_arrive(&bar_tempty[a]);
      if (++a == 2) { a = 0; aph ^= 1; }
    }
    if (epi_tid == 0) bulk_wait_all<0>();
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
