// Attention cores of the transformer layers.
//
//   attn_global_kernel  SelfAttention over the whole 250-frame window (model.py:241-257, 364-366):
//                       S = Q K^T and O = P V on tcgen05 (UMMA 128x256x16 and 128x64x16), S and O in
//                       TMEM, operands TMA-staged (Q, K, V^T) or written by the softmax threads (P).
//   attn_local_tc_kernel LocalSelfAttention (model.py:409-471) as one banded attention on tcgen05 (see below).
//
// Both read RoPE-ready bf16 projections produced by the GEMM epilogues and write the bf16 operand of the
// output projection.  Sequence rows are padded 250 -> 256 per window (row = b * 256 + t).
#pragma once
#include "ptx.cuh"

namespace a2m {

#ifdef A2M_FFN_TIMING
#define AT_STAMP(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) g_ffn_timing[(i)] = clock64(); } while (0)
#else
#define AT_STAMP(i) do { } while (0)
#endif

constexpr int ATT_T = 250;    // real frames per window
constexpr int ATT_TP = 256;   // padded rows per window
constexpr int ATT_HD = 64;    // head dim
constexpr int ATT_HEADS = 4;

// 2^x, flush-to-zero, one MUFU op (exp2f() adds a denormal-range rescale around it)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16x2_att(float lo, float hi) {
  __nv_bfloat162 v = op2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------ global
constexpr int AG_THREADS = 256;           // 8 warps: TMEM quadrant = warp & 3, key half = warp >> 2
constexpr int AG_SQ = 128 * 64 * 2;       // 16 KB
constexpr int AG_SK = 256 * 64 * 2;       // 32 KB
constexpr int AG_SV = 256 * 64 * 2;       // 32 KB: V[key][d], MN-major B operand of P.V
constexpr int AG_SP = 4 * 128 * 64 * 2;   // 64 KB: 4 k-blocks of P [128 q x 64 keys]; ALIASES Q and K (dead after S)
constexpr size_t AG_SMEM = 1024 + AG_SP + AG_SV + 4 * 128 * 4 + 128;   // ~100 KB -> two CTAs per SM
constexpr uint32_t AG_TMEM_COLS = 256;    // S: 256 columns; O re-uses columns 0..63 once S has been consumed

// grid = (2 m-tiles, heads, B).  tmQ: Q  [B*256, ldq] box {64,128}; tmK: K [B*256, ldkv] box {64,256};
// tmV: V [B*256, ldkv] (columns 256 + h*64 ..) box {64,256}.  Q and K arrive RoPE-rotated (GEMM epilogue).
__global__ void __launch_bounds__(AG_THREADS, 2)
attn_global_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ O, int ldo, int v_col0,
                   float* __restrict__ lse, const DropParams* __restrict__ drop, uint32_t drop_site) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AG_SQ;
  uint8_t* sP = smem;            // overwrites Q/K after the S MMAs have completed
  uint8_t* sV = smem + AG_SP;
  float* sMax = reinterpret_cast<float*>(sV + AG_SV);   // [2 key halves][128 rows]
  float* sSum = sMax + 2 * 128;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sSum + 2 * 128);
  uint64_t* bar_s = bar_load + 1;
  uint64_t* bar_o = bar_load + 2;
  uint64_t* bar_v = bar_load + 3;       // V on its own barrier: the S MMAs start as soon as Q and K have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 4);

  const int mt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, kh = warp >> 2;   // this thread: query row quad * 32 + lane, keys kh * 128 .. + 127

  pdl_launch_dependents();
  AT_STAMP(64);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_v, 1);
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
  AT_STAMP(65);

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, AG_SQ + AG_SK);
    tma_load_2d(sQ, &tmQ, bar_load, h * ATT_HD, b * ATT_TP + mt * 128);
    tma_load_2d(sK, &tmK, bar_load, h * ATT_HD, b * ATT_TP);
    mbar_arrive_expect_tx(bar_v, AG_SV);
    tma_load_2d(sV, &tmV, bar_v, v_col0 + h * ATT_HD, b * ATT_TP);
    mbar_wait(bar_load, 0);
    tc_fence_after();
    AT_STAMP(66);
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
  AT_STAMP(67);
  const int row = quad * 32 + lane;  // query row inside the tile
  const uint32_t t_row = (static_cast<uint32_t>(quad * 32) << 16);
  // softmax((q / 8) . k): scale folded into the exponent; exp2 with log2(e) pre-multiplied
  const float kscale = 0.125f * 1.4426950408889634f;
  // training: dropout of the attention weights (model.py:254-255), after the normalisation
  const uint32_t dthresh = drop ? drop->thresh : 0u;
  const float dinv = drop ? drop->inv_keep : 1.f;
  const uint32_t dkey = drop ? drop_key(drop->seed, drop_site) : 0u;
  const uint32_t dbase = ((static_cast<uint32_t>(b) * ATT_HEADS + h) * ATT_TP + mt * 128 + row) * ATT_TP;
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int key0 = kh * 128 + c * 32;
    uint32_t r[32];
    tmem_ld_x32(tmem_S + t_row + key0, r);
    tmem_ld_wait();
    if (key0 + 32 <= ATT_T) {          // every chunk but the last one of the upper key half: no padded keys, no per-element test
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (key0 + j < ATT_T) mx = fmaxf(mx, __uint_as_float(r[j]));
    }
  }
  // the two threads of a row exchange their half-row maxima (keys 0..127 always hold real keys: never -inf)
  sMax[kh * 128 + row] = mx;
  __syncthreads();
  mx = fmaxf(sMax[row], sMax[128 + row]);
  const float mxk = mx * kscale;
  float sum = 0.f;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int key0 = kh * 128 + c * 32;
    uint32_t r[32];
    tmem_ld_x32(tmem_S + t_row + key0, r);
    tmem_ld_wait();
    float p[32];
    if (key0 + 32 <= ATT_T) {
#pragma unroll
      for (int j = 0; j < 32; ++j) p[j] = ex2_ftz(fmaf(__uint_as_float(r[j]), kscale, -mxk));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float e = ex2_ftz(fmaf(__uint_as_float(r[j]), kscale, -mxk));
        p[j] = (key0 + j < ATT_T) ? e : 0.f;  // padded keys 250..255 are masked out
      }
    }
    // P (bf16) into the K-major 128B-swizzled A-operand layout: k-block = key / 64
    uint8_t* pb = sP + (key0 >> 6) * (128 * 64 * 2);
    const int colb = key0 & 63;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __nv_bfloat162 h0 = op2_rn(p[8 * q], p[8 * q + 1]);
      __nv_bfloat162 h1 = op2_rn(p[8 * q + 2], p[8 * q + 3]);
      __nv_bfloat162 h2 = op2_rn(p[8 * q + 4], p[8 * q + 5]);
      __nv_bfloat162 h3 = op2_rn(p[8 * q + 6], p[8 * q + 7]);
      // the row sum must match what the tensor core will see: accumulate the ROUNDED probabilities
      sum += (op2_sum(h0) + op2_sum(h1)) + (op2_sum(h2) + op2_sum(h3));
      if (dthresh != 0u) {   // the P.V operand is the dropped-out weights; the normaliser above is not
        const uint32_t i0 = dbase + key0 + 8 * q;
        h0 = op2_rn(p[8 * q] * drop_mul(dkey, i0, dthresh, dinv), p[8 * q + 1] * drop_mul(dkey, i0 + 1, dthresh, dinv));
        h1 = op2_rn(p[8 * q + 2] * drop_mul(dkey, i0 + 2, dthresh, dinv), p[8 * q + 3] * drop_mul(dkey, i0 + 3, dthresh, dinv));
        h2 = op2_rn(p[8 * q + 4] * drop_mul(dkey, i0 + 4, dthresh, dinv), p[8 * q + 5] * drop_mul(dkey, i0 + 5, dthresh, dinv));
        h3 = op2_rn(p[8 * q + 6] * drop_mul(dkey, i0 + 6, dthresh, dinv), p[8 * q + 7] * drop_mul(dkey, i0 + 7, dthresh, dinv));
      }
      uint4 v;
      v.x = *reinterpret_cast<uint32_t*>(&h0);
      v.y = *reinterpret_cast<uint32_t*>(&h1);
      v.z = *reinterpret_cast<uint32_t*>(&h2);
      v.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(pb + sw128_offset(row, colb + 8 * q)) = v;
    }
  }
  sSum[kh * 128 + row] = sum;
  AT_STAMP(68);
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to tcgen05.mma
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    mbar_wait(bar_v, 0);
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
  AT_STAMP(69);
  const float inv = __fdividef(1.0f, sSum[row] + sSum[128 + row]);
  // training: natural-log softmax denominator per (row, head), P = exp(s / 8 - lse)  (attention_bwd.cuh)
  if (lse != nullptr && kh == 0)
    lse[static_cast<size_t>(b * ATT_TP + mt * 128 + row) * ATT_HEADS + h] = mx * 0.125f + __logf(sSum[row] + sSum[128 + row]);
  // head-dim half kh of this row: 32 bf16 = 64 bytes
  __nv_bfloat16* dst = O + static_cast<size_t>(b * ATT_TP + mt * 128 + row) * ldo + h * ATT_HD + kh * 32;
  {
    uint32_t r[32];
    tmem_ld_x32(tmem_O + t_row + kh * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 v;
      __nv_bfloat162 h0 = op2_rn(__uint_as_float(r[8 * q]) * inv, __uint_as_float(r[8 * q + 1]) * inv);
      __nv_bfloat162 h1 = op2_rn(__uint_as_float(r[8 * q + 2]) * inv, __uint_as_float(r[8 * q + 3]) * inv);
      __nv_bfloat162 h2 = op2_rn(__uint_as_float(r[8 * q + 4]) * inv, __uint_as_float(r[8 * q + 5]) * inv);
      __nv_bfloat162 h3 = op2_rn(__uint_as_float(r[8 * q + 6]) * inv, __uint_as_float(r[8 * q + 7]) * inv);
      v.x = *reinterpret_cast<uint32_t*>(&h0);
      v.y = *reinterpret_cast<uint32_t*>(&h1);
      v.z = *reinterpret_cast<uint32_t*>(&h2);
      v.w = *reinterpret_cast<uint32_t*>(&h3);
      reinterpret_cast<uint4*>(dst)[q] = v;
    }
  }
  AT_STAMP(70);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<AG_TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ local (tensor cores)
// LocalSelfAttention (model.py:409-471) restated for the tensor cores.
//
// Reference: pad the normalised sequence with 3 zero rows left / 3 right (256 padded rows), attend inside 31
// windows of 16 padded rows at stride 8 with RoPE positions 0..15 INSIDE each window, scatter-add window row r
// to output row (start + r) of the unpadded buffer (so output row j holds the result for padded row j = token
// j-3; rows >= 250 are dropped) and divide by the number of windows that covered the row.
//
// Two identities make this one banded attention:
//  (1) RoPE scores depend only on the position DIFFERENCE: <R(a) q, R(b) k> = <R(a+c) q, R(b+c) k>.  Rotating
//      q and k once with their absolute row index (done in the projection GEMM epilogue, as for the global
//      layers) gives the same logits as the per-window positions 0..15, up to fp32 rounding of the table.
//  (2) the output projection is linear and bias free, so the mean over the <= 2 covering windows commutes with
//      it: P = mean over windows of that window's softmax row, then one P.V product.
// Zero-padding tokens (token < 0 or >= 250) have q = k = v = 0 (bias-free projections) but still own a softmax
// slot with logit 0.  They are produced by TMA out-of-bounds zero fill: the tensor maps view each window as
// [250 rows] so negative rows and rows 250.. come back as zeros.
//
// Per CTA: 128 padded query rows j = 128 mt .. +127 of one (window b, head h); keys = padded rows 128 mt - 8 ..
// 128 mt + 135 (144 rows).  S = Q K^T (UMMA 128x144x16) in TMEM, banded two-window softmax on CUDA cores straight
// from TMEM, P (bf16) into the swizzled A-operand layout, O = P V (UMMA 128x64x16, V MN-major), bf16 store.
constexpr int AL_THREADS = 128;
constexpr int AL_NK = 144;                 // keys per tile
constexpr int AL_SQ = 128 * 64 * 2;        // 16 KB
constexpr int AL_SK = AL_NK * 64 * 2;      // 18 KB
constexpr int AL_SV = AL_NK * 64 * 2;      // 18 KB, V[key][d]
constexpr int AL_SP = 3 * 128 * 64 * 2;    // 48 KB: key blocks 0-63, 64-127, 128-143
constexpr size_t AL_SMEM = 1024 + AL_SQ + AL_SK + AL_SV + AL_SP + 128;  // ~101 KB -> two CTAs per SM
constexpr uint32_t AL_TMEM_COLS = 256;     // S: columns 0..143, O: columns 192..255

// tmQ / tmK / tmV are 3-D maps {cols, 250 rows, B} over the q||c and k||v projection buffers.
__global__ void __launch_bounds__(AL_THREADS, 2)
attn_local_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* O, int ldo, int v_col0,
                     const DropParams* __restrict__ drop, uint32_t drop_site) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AL_SQ;
  uint8_t* sV = sK + AL_SK;
  uint8_t* sP = sV + AL_SV;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sP + AL_SP);
  uint64_t* bar_s = bar_load + 1;
  uint64_t* bar_o = bar_load + 2;
  uint64_t* bar_v = bar_load + 3;       // V on its own barrier: the S MMAs start as soon as Q and K have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 4);

  const int mt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_launch_dependents();
  AT_STAMP(72);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_v, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<AL_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 192;
  pdl_wait();
  AT_STAMP(73);

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, AL_SQ + AL_SK);
    // padded row j <-> token j - 3; rows outside [0, 250) are zero-filled by the TMA unit
    tma_load_3d(sQ, &tmQ, bar_load, h * ATT_HD, mt * 128 - 3, b);
    tma_load_3d(sK, &tmK, bar_load, h * ATT_HD, mt * 128 - 8 - 3, b);
    mbar_arrive_expect_tx(bar_v, AL_SV);
    tma_load_3d(sV, &tmV, bar_v, v_col0 + h * ATT_HD, mt * 128 - 8 - 3, b);
    mbar_wait(bar_load, 0);
    tc_fence_after();
    AT_STAMP(74);
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, AL_NK);
    const uint64_t dq = umma_desc_sw128(smem_u32(sQ));
    const uint64_t dk = umma_desc_sw128(smem_u32(sK));
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem_S, umma_desc_advance_k(dq, k * 32), umma_desc_advance_k(dk, k * 32), idesc_s, k != 0 ? 1u : 0u);
    umma_commit(bar_s);
  }
  __syncwarp();

  mbar_wait(bar_s, 0);
  tc_fence_after();
  AT_STAMP(75);
  const int row = warp * 32 + lane;        // query row in the tile
  const int j = mt * 128 + row;            // padded row == output row
  const uint32_t t_row = static_cast<uint32_t>(warp * 32) << 16;
  // key columns (tile relative) of this row's two windows:  A = window floor(j/8)-1 : [off, off+16),
  //                                                           B = window floor(j/8)   : [off+8, off+24)
  // relative to the warp's 48-column slab that starts at column 32 * warp.
  const int off = (lane >> 3) << 3;
  const bool has_a = j >= 8;               // window index >= 0
  const bool has_b = j < 248;              // window index <= 30
  float s[48];
  {
    uint32_t r0[32], r1[16];
    tmem_ld_x32(tmem_S + t_row + warp * 32, r0);
    tmem_ld_x16(tmem_S + t_row + warp * 32 + 32, r1);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c) s[c] = __uint_as_float(r0[c]);
#pragma unroll
    for (int c = 0; c < 16; ++c) s[32 + c] = __uint_as_float(r1[c]);
  }
  const float kscale = 0.125f * 1.4426950408889634f;   // (q / sqrt(64)) . k, exp2 domain
  float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
  for (int k = 0; k < 6; ++k) {   // per 8-column chunk, then the two chunks of each window
    float cm = s[8 * k];
#pragma unroll
    for (int i = 1; i < 8; ++i) cm = fmaxf(cm, s[8 * k + i]);
    const int rk = k - (off >> 3);
    if (rk == 0 || rk == 1) ma = fmaxf(ma, cm);
    if (rk == 1 || rk == 2) mb = fmaxf(mb, cm);
  }
  const float wn = (has_a && has_b) ? 0.5f : 1.0f;      // divide by the number of covering windows
  const uint32_t dthresh = drop ? drop->thresh : 0u;
  if (dthresh == 0u) {
    // Inference path.  Window A covers the 8-column chunks off/8 and off/8 + 1 of the slab, window B off/8 + 1 and
    // off/8 + 2: every element is exponentiated once per window with the exponent clamped to <= 0 (in-band elements
    // never exceed their window's maximum, out-of-band ones are multiplied by a zero chunk weight afterwards), so the
    // per-lane band position only enters through six per-chunk weights instead of per-element predicates.
    const int k0 = off >> 3;
    const float mak = ma * kscale, mbk = mb * kscale;
    float suma = 0.f, sumb = 0.f;
    float ea[48];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const bool ua = (k == k0) || (k == k0 + 1), ub = (k == k0 + 1) || (k == k0 + 2);
      float ca = 0.f, cb = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = s[8 * k + i] * kscale;
        const float a = ex2_ftz(fminf(x - mak, 0.f)), bb = ex2_ftz(fminf(x - mbk, 0.f));
        ea[8 * k + i] = a;
        s[8 * k + i] = bb;
        ca += a;
        cb += bb;
      }
      if (ua) suma += ca;
      if (ub) sumb += cb;
    }
    const float ia = has_a ? __fdividef(wn, suma) : 0.f;
    const float ib = has_b ? __fdividef(wn, sumb) : 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float wa_k = ((k == k0) || (k == k0 + 1)) ? ia : 0.f;
      const float wb_k = ((k == k0 + 1) || (k == k0 + 2)) ? ib : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s[8 * k + i] = fmaf(ea[8 * k + i], wa_k, s[8 * k + i] * wb_k);
    }
  } else {
    float suma = 0.f, sumb = 0.f;
  #pragma unroll
    for (int c = 0; c < 48; ++c) {
      const bool in_a = static_cast<unsigned>(c - off) < 16u;
      const bool in_b = static_cast<unsigned>(c - off - 8) < 16u;
      if (in_a) suma += exp2f((s[c] - ma) * kscale);
      if (in_b) sumb += exp2f((s[c] - mb) * kscale);
    }
    const float ia = has_a ? __fdividef(wn, suma) : 0.f;
    const float ib = has_b ? __fdividef(wn, sumb) : 0.f;
    // training: each window's attention weights are dropped out independently (model.py:443 -> 254-255);
    // element index ((b, h, window, row in window), key in window)
    const float dinv = drop ? drop->inv_keep : 1.f;
    const uint32_t dkey = drop ? drop_key(drop->seed, drop_site) : 0u;
    const int wa = (j >> 3) - 1, wb = j >> 3;
    const uint32_t bh = static_cast<uint32_t>(b) * ATT_HEADS + h;
    const uint32_t da0 = ((bh * 31u + static_cast<uint32_t>(wa)) * 16u + static_cast<uint32_t>(j - 8 * wa)) * 16u;
    const uint32_t db0 = ((bh * 31u + static_cast<uint32_t>(wb)) * 16u + static_cast<uint32_t>(j - 8 * wb)) * 16u;
  #pragma unroll
    for (int c = 0; c < 48; ++c) {
      const bool in_a = static_cast<unsigned>(c - off) < 16u;
      const bool in_b = static_cast<unsigned>(c - off - 8) < 16u;
      float p = 0.f;
      if (in_a) {
        float pa = exp2f((s[c] - ma) * kscale) * ia;
        if (dthresh != 0u) pa *= drop_mul(dkey, da0 + static_cast<uint32_t>(c - off), dthresh, dinv);
        p += pa;
      }
      if (in_b) {
        float pb = exp2f((s[c] - mb) * kscale) * ib;
        if (dthresh != 0u) pb *= drop_mul(dkey, db0 + static_cast<uint32_t>(c - off - 8), dthresh, dinv);
        p += pb;
      }
      s[c] = p;
    }
  }
  // P row (144 keys = 18 chunks of 8) into the swizzled K-major layout; this warp's slab is chunks 4w .. 4w+5
  {
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int cc = 0; cc < 18; ++cc) {
      if (cc < 4 * warp || cc >= 4 * warp + 6)
        *reinterpret_cast<uint4*>(sP + (cc >> 3) * (128 * 128) + sw128_offset(row, (cc & 7) * 8)) = zero;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int cc = 4 * warp + i;
      uint4 v;
      v.x = pack_bf16x2_att(s[8 * i], s[8 * i + 1]);
      v.y = pack_bf16x2_att(s[8 * i + 2], s[8 * i + 3]);
      v.z = pack_bf16x2_att(s[8 * i + 4], s[8 * i + 5]);
      v.w = pack_bf16x2_att(s[8 * i + 6], s[8 * i + 7]);
      *reinterpret_cast<uint4*>(sP + (cc >> 3) * (128 * 128) + sw128_offset(row, (cc & 7) * 8)) = v;
    }
  }
  AT_STAMP(76);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    mbar_wait(bar_v, 0);
    tc_fence_after();
    constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, 64);
    const uint64_t dv = umma_desc_sw128(smem_u32(sV));
#pragma unroll
    for (int ks = 0; ks < AL_NK / 16; ++ks) {   // 9 key steps of 16
      const uint64_t dp = umma_desc_sw128(smem_u32(sP + (ks >> 2) * (128 * 128)));
      umma_bf16(tmem_O, umma_desc_advance_k(dp, (ks & 3) * 32), umma_desc_advance_k(dv, ks * 2048), idesc_o, ks != 0 ? 1u : 0u);
    }
    umma_commit(bar_o);
  }
  __syncwarp();

  mbar_wait(bar_o, 0);
  tc_fence_after();
  AT_STAMP(77);
  __nv_bfloat16* dst = O + static_cast<size_t>(b * ATT_TP + j) * ldo + h * ATT_HD;
  const bool real = j < ATT_T;   // rows 250..255 are written as zeros (kept finite for the padded projections)
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld_x32(tmem_O + t_row + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 v;
      v.x = real ? pack_bf16x2_att(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])) : 0u;
      v.y = real ? pack_bf16x2_att(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])) : 0u;
      v.z = real ? pack_bf16x2_att(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])) : 0u;
      v.w = real ? pack_bf16x2_att(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])) : 0u;
      reinterpret_cast<uint4*>(dst + c * 32)[q] = v;
    }
  }
  AT_STAMP(78);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<AL_TMEM_COLS>(tmem_base);
  }
}

}  // namespace a2m
