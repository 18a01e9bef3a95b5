// Attention cores of the transformer layers.
//
//   attn_global_kernel  SelfAttention over the whole 250-frame window (model.py:241-257, 364-366):
//                       S = Q K^T and O = P V on tcgen05 (UMMA 128x256x16 and 128x64x16), S and O in
//                       TMEM, operands TMA-staged (Q, K, V^T) or written by the softmax threads (P).
//   attn_local_kernel   LocalSelfAttention (model.py:409-471) on CUDA cores: 16-frame windows at
//                       stride 8; tiny contractions (16x16x64), latency bound.
//
// Both read RoPE-ready bf16 projections produced by the GEMM epilogues and write the bf16 operand of the
// output projection.  Sequence rows are padded 250 -> 256 per window (row = b * 256 + t).
#pragma once
#include "ptx.cuh"

namespace a2m {

constexpr int ATT_T = 250;    // real frames per window
constexpr int ATT_TP = 256;   // padded rows per window
constexpr int ATT_HD = 64;    // head dim
constexpr int ATT_HEADS = 4;

// ------------------------------------------------------------------------------------------ global
constexpr int AG_THREADS = 128;
constexpr int AG_SQ = 128 * 64 * 2;       // 16 KB
constexpr int AG_SK = 256 * 64 * 2;       // 32 KB
constexpr int AG_SV = 256 * 64 * 2;       // 32 KB: V[key][d], MN-major B operand of P.V
constexpr int AG_SP = 4 * 128 * 64 * 2;   // 64 KB: 4 k-blocks of P [128 q x 64 keys]; ALIASES Q and K (dead after S)
constexpr size_t AG_SMEM = 1024 + AG_SP + AG_SV + 128;   // ~98 KB -> two CTAs per SM
constexpr uint32_t AG_TMEM_COLS = 256;    // S: 256 columns; O re-uses columns 0..63 once S has been consumed

// grid = (2 m-tiles, heads, B).  tmQ: Q  [B*256, ldq] box {64,128}; tmK: K [B*256, ldkv] box {64,256};
// tmV: V [B*256, ldkv] (columns 256 + h*64 ..) box {64,256}.  Q and K arrive RoPE-rotated (GEMM epilogue).
__global__ void __launch_bounds__(AG_THREADS, 2)
attn_global_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ O, int ldo, int v_col0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AG_SQ;
  uint8_t* sP = smem;            // overwrites Q/K after the S MMAs have completed
  uint8_t* sV = smem + AG_SP;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sV + AG_SV);
  uint64_t* bar_s = bar_load + 1;
  uint64_t* bar_o = bar_load + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 3);

  const int mt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<AG_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;  // 256 columns
  const uint32_t tmem_O = tmem_base;  // 64 columns, written only after every thread has read S
  pdl_wait();

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, AG_SQ + AG_SK + AG_SV);
    tma_load_2d(sQ, &tmQ, bar_load, h * ATT_HD, b * ATT_TP + mt * 128);
    tma_load_2d(sK, &tmK, bar_load, h * ATT_HD, b * ATT_TP);
    tma_load_2d(sV, &tmV, bar_load, v_col0 + h * ATT_HD, b * ATT_TP);
    mbar_wait(bar_load, 0);
    tc_fence_after();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256);
    const uint64_t dq = umma_desc_sw128(smem_u32(sQ));
    const uint64_t dk = umma_desc_sw128(smem_u32(sK));
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem_S, umma_desc_advance_k(dq, k * 32), umma_desc_advance_k(dk, k * 32), idesc_s, k != 0 ? 1u : 0u);
    umma_commit(bar_s);
  }
  __syncwarp();

  // ---- softmax over keys: one query row per thread (TMEM lane = row) ----
  mbar_wait(bar_s, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;  // query row inside the tile
  const uint32_t t_row = (static_cast<uint32_t>(warp * 32) << 16);
  // softmax((q / 8) . k): scale folded into the exponent; exp2 with log2(e) pre-multiplied
  const float kscale = 0.125f * 1.4426950408889634f;
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    uint32_t r[32];
    tmem_ld_x32(tmem_S + t_row + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c * 32 + j < ATT_T) mx = fmaxf(mx, __uint_as_float(r[j]));
  }
  float sum = 0.f;
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    uint32_t r[32];
    tmem_ld_x32(tmem_S + t_row + c * 32, r);
    tmem_ld_wait();
    float p[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float e = exp2f((__uint_as_float(r[j]) - mx) * kscale);
      p[j] = (c * 32 + j < ATT_T) ? e : 0.f;  // padded keys 250..255 are masked out
    }
    // P (bf16) into the K-major 128B-swizzled A-operand layout: k-block = key / 64
    uint8_t* pb = sP + (c >> 1) * (128 * 64 * 2);
    const int colb = (c & 1) * 32;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(p[8 * q], p[8 * q + 1]);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(p[8 * q + 2], p[8 * q + 3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(p[8 * q + 4], p[8 * q + 5]);
      __nv_bfloat162 h3 = __floats2bfloat162_rn(p[8 * q + 6], p[8 * q + 7]);
      // the row sum must match what the tensor core will see: accumulate the ROUNDED probabilities
      sum += __low2float(h0) + __high2float(h0) + __low2float(h1) + __high2float(h1) + __low2float(h2) +
             __high2float(h2) + __low2float(h3) + __high2float(h3);
      uint4 v;
      v.x = *reinterpret_cast<uint32_t*>(&h0);
      v.y = *reinterpret_cast<uint32_t*>(&h1);
      v.z = *reinterpret_cast<uint32_t*>(&h2);
      v.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(pb + sw128_offset(row, colb + 8 * q)) = v;
    }
  }
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to tcgen05.mma
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after();
    constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, 64);  // B = V[key][d]: MN-major
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      const uint64_t dp = umma_desc_sw128(smem_u32(sP + kb * (128 * 64 * 2)));
      const uint64_t dv = umma_desc_sw128(smem_u32(sV + kb * (64 * 128)));
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_O, umma_desc_advance_k(dp, k * 32), umma_desc_advance_k(dv, k * 2048), idesc_o,
                  (kb | k) != 0 ? 1u : 0u);
    }
    umma_commit(bar_o);
  }
  __syncwarp();

  mbar_wait(bar_o, 0);
  tc_fence_after();
  const float inv = __fdividef(1.0f, sum);
  __nv_bfloat16* dst = O + static_cast<size_t>(b * ATT_TP + mt * 128 + row) * ldo + h * ATT_HD;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld_x32(tmem_O + t_row + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 v;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(r[8 * q]) * inv, __uint_as_float(r[8 * q + 1]) * inv);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(r[8 * q + 2]) * inv, __uint_as_float(r[8 * q + 3]) * inv);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[8 * q + 4]) * inv, __uint_as_float(r[8 * q + 5]) * inv);
      __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(r[8 * q + 6]) * inv, __uint_as_float(r[8 * q + 7]) * inv);
      v.x = *reinterpret_cast<uint32_t*>(&h0);
      v.y = *reinterpret_cast<uint32_t*>(&h1);
      v.z = *reinterpret_cast<uint32_t*>(&h2);
      v.w = *reinterpret_cast<uint32_t*>(&h3);
      reinterpret_cast<uint4*>(dst + c * 32)[q] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<AG_TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ local
// One warp per (window b, head h, block w of 8 output rows j = 8w .. 8w+7).  The reference pads the
// normalised sequence with 3 zero rows on the left (and 3 on the right), runs attention in 31 windows
// of 16 padded rows at stride 8, and scatter-adds window row r into output row (start + r) of the
// UNPADDED buffer (model.py:422-469).  Hence output row j = mean over the windows {w-1, w} that exist
// of the attention result for padded row j = token j-3; zero-padding tokens have q = k = v = 0 (the
// projections are bias free) but still occupy a slot in each softmax.  RoPE positions are the row
// index inside the window (rope.py:40-41), so K is rotated per window.
// The output projection is linear and bias free, so it is applied AFTER the mean (one GEMM on 256 rows).
constexpr int AL_WARPS = 4;
constexpr int AL_SK_STRIDE = 66;  // floats; conflict-free 64-bit reads for 16 keys
struct AlSmem {
  float k[16][AL_SK_STRIDE];
  float q[8][64];
  float p[8][16];
};

__global__ void __launch_bounds__(AL_WARPS * 32)
attn_local_kernel(const __nv_bfloat16* Q, int ldq, const __nv_bfloat16* K, const __nv_bfloat16* V, int ldkv,
                  __nv_bfloat16* O, int ldo,
                  const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int total_warps) {
  __shared__ AlSmem sm_all[AL_WARPS];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int id = blockIdx.x * AL_WARPS + warp;
  if (id >= total_warps) return;
  AlSmem& sm = sm_all[warp];
  const int w = id & 31;
  const int h = (id >> 5) & 3;
  const int b = id >> 7;
  const size_t rowbase = static_cast<size_t>(b) * ATT_TP;

  float o[8][2];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) o[jj][0] = o[jj][1] = 0.f;
  int count = 0;

#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    const int win = w - 1 + which;  // window w-1 first, then window w
    if (win < 0 || win > 30) continue;
    ++count;
    const int s = 8 * win;          // first padded row of the window
    float v[16][2];
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const int tok = s + kk - 3;
      float k0 = 0.f, k1 = 0.f;
      v[kk][0] = v[kk][1] = 0.f;
      if (tok >= 0 && tok < ATT_T) {
        const size_t r = rowbase + tok;
        const __nv_bfloat162 kv = *reinterpret_cast<const __nv_bfloat162*>(K + r * ldkv + h * ATT_HD + 2 * lane);
        const __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(V + r * ldkv + h * ATT_HD + 2 * lane);
        k0 = __low2float(kv); k1 = __high2float(kv);
        v[kk][0] = __low2float(vv); v[kk][1] = __high2float(vv);
      }
      const float c = __ldg(rope_cos + kk * 32 + lane), sn = __ldg(rope_sin + kk * 32 + lane);
      *reinterpret_cast<float2*>(&sm.k[kk][2 * lane]) = make_float2(k0 * c - k1 * sn, k0 * sn + k1 * c);
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = 8 * w + jj;      // output row == padded row
      const int tok = j - 3;
      const int pos = j - s;         // position inside the window (0..15)
      float q0 = 0.f, q1 = 0.f;
      if (tok >= 0 && tok < ATT_T) {
        const __nv_bfloat162 qv =
            *reinterpret_cast<const __nv_bfloat162*>(Q + (rowbase + tok) * ldq + h * ATT_HD + 2 * lane);
        q0 = __low2float(qv); q1 = __high2float(qv);
      }
      const float c = __ldg(rope_cos + pos * 32 + lane), sn = __ldg(rope_sin + pos * 32 + lane);
      *reinterpret_cast<float2*>(&sm.q[jj][2 * lane]) = make_float2(q0 * c - q1 * sn, q0 * sn + q1 * c);
    }
    __syncwarp();

    // scores: lane -> key kk = lane & 15, queries jj = 4 * (lane >> 4) + u
    const int kk = lane & 15, g = lane >> 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int d = 0; d < 64; d += 2) {
      const float2 kv = *reinterpret_cast<const float2*>(&sm.k[kk][d]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 qv = *reinterpret_cast<const float2*>(&sm.q[4 * g + u][d]);
        acc[u] = fmaf(qv.x, kv.x, acc[u]);
        acc[u] = fmaf(qv.y, kv.y, acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float sc = acc[u] * 0.125f;  // query / sqrt(64)
      float m = sc;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
      const float e = __expf(sc - m);
      float t = e;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
      sm.p[4 * g + u][kk] = __fdividef(e, t);
    }
    __syncwarp();

    // O += P V : lane owns head dims (2 lane, 2 lane + 1)
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const float pw = sm.p[jj][k2];
        o[jj][0] = fmaf(pw, v[k2][0], o[jj][0]);
        o[jj][1] = fmaf(pw, v[k2][1], o[jj][1]);
      }
    }
    __syncwarp();
  }

  const float inv = 1.0f / static_cast<float>(count);
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = 8 * w + jj;
    // pad rows 250..255 are written as zeros: they feed the output projection of the padded rows and must
    // stay finite (stale memory there could be NaN/Inf)
    const bool real = j < ATT_T;
    *reinterpret_cast<__nv_bfloat162*>(O + (rowbase + j) * ldo + h * ATT_HD + 2 * lane) =
        __floats2bfloat162_rn(real ? o[jj][0] * inv : 0.f, real ? o[jj][1] * inv : 0.f);
  }
}

}  // namespace a2m
