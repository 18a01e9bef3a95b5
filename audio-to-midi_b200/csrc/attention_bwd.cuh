// Backward of the attention cores (the `eqx.filter_value_and_grad` of model.py:241-257 / 409-471 inside train.py:50).
//
//   attn_global_bwd_kernel  one CTA per (window b, head h); all five products of the flash-style backward on
//                           tcgen05 with fp32 accumulators in TMEM:
//                               S = Q K^T,  dP = dO V^T                      (recomputed, K-major operands)
//                               P = exp(S/8 - lse),  dS = P o (dP - D) / 8   (CUDA cores, from TMEM)
//                               dV += P^T dO,  dK += dS^T Q                  (A operand MN-major: P / dS as stored)
//                               dQ += dS K                                   (B operand MN-major: K as stored)
//                           then the inverse RoPE rotation of dQ / dK in the epilogue, bf16 stores.
//
// Inputs are the RoPE-rotated bf16 projections the forward wrote (q||c and k||v buffers), the forward output O,
// the softmax log-sum-exp per (row, head), and dO (bf16).  Outputs: dQ (raw, pre-RoPE) into the q columns of the
// dQC buffer, dK (raw) || dV into the dKV buffer.  Rows 250..255 of every window are written as zeros.
#pragma once
#include "attention.cuh"
#include "gemm_wgrad.cuh"

namespace a2m {

constexpr int AGB_THREADS = 256;
constexpr int AGBG_THREADS = 512;            // attn_global_bwd_kernel: four warps per TMEM quadrant (32 key columns / 16 accumulator columns each)
constexpr int AGB_TILE = 128 * 64 * 2;               // 16 KB: one 128-row tile of a [rows][64] bf16 operand
constexpr int AGB_OPER = 2 * AGB_TILE;               // 32 KB: 256 rows
constexpr int AGB_PS = 2 * 128 * 64 * 2;             // 32 KB: P or dS tile [128 q][128 keys] as two k-blocks of 64 keys
constexpr size_t AGB_SMEM = 1024 + 4 * AGB_OPER + 2 * AGB_PS + 128;   // ~193 KB
constexpr uint32_t AGB_TMEM_COLS = 512;

// TMEM columns
constexpr uint32_t AGB_C_S = 0, AGB_C_DP = 128, AGB_C_DK = 256, AGB_C_DV = 320, AGB_C_DQ0 = 384, AGB_C_DQ1 = 448;

// tmQ: q||c buffer [B*256, ldq], tmK / tmV: k||v buffer, tmDO: dO [B*256, 256]; all box {64, 256}.
// lse [B*256, 4] fp32 (natural log of the row's softmax denominator, including the running max);
// O, dO row-major bf16 with leading dimension ldo.
__global__ void __launch_bounds__(AGBG_THREADS, 1)
attn_global_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                       const __nv_bfloat16* __restrict__ O, const __nv_bfloat16* __restrict__ dO, int ldo,
                       const float* __restrict__ lse, const float* __restrict__ rope_cos, const float* __restrict__ rope_sin,
                       __nv_bfloat16* __restrict__ dQ, int lddq, __nv_bfloat16* __restrict__ dKV, int lddkv, int v_col0,
                       const DropParams* __restrict__ drop, uint32_t drop_site) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AGB_OPER;
  uint8_t* sV = sK + AGB_OPER;
  uint8_t* sDO = sV + AGB_OPER;
  uint8_t* sP = sDO + AGB_OPER;
  uint8_t* sDS = sP + AGB_PS;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sDS + AGB_PS);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);

  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, qtr = warp >> 2;    // TMEM lane quadrant; which 32 of a tile's 128 key columns (16 of an accumulator's 64)
  const int row = quad * 32 + lane;              // row inside a 128-row tile
  const uint32_t t_row = static_cast<uint32_t>(quad * 32) << 16;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<AGB_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // RoPE rows of this thread's two output rows (row, 128 + row), its 8 pairs: constants, fetched before the dependency wait
  float4 pre_c[2][2], pre_s[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int pos = min(i * 128 + row, ATT_T - 1);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      pre_c[i][j] = __ldg(reinterpret_cast<const float4*>(rope_cos + pos * 32 + qtr * 8) + j);
      pre_s[i][j] = __ldg(reinterpret_cast<const float4*>(rope_sin + pos * 32 + qtr * 8) + j);
    }
  }
  pdl_wait();

  constexpr uint32_t idesc_kk = umma_idesc_bf16(128, 128);        // S, dP: both K-major
  constexpr uint32_t idesc_ab = umma_idesc_bf16_abmn(128, 64);    // dV, dK: A (P / dS) and B (dO / Q) MN-major
  constexpr uint32_t idesc_b = umma_idesc_bf16_bmn(128, 64);      // dQ: A (dS) K-major, B (K) MN-major

  auto issue_scores = [&](int i, int j) {   // S = Q_i K_j^T, dP = dO_i V_j^T
    const uint64_t dq = umma_desc_sw128(smem_u32(sQ + i * AGB_TILE)), dk = umma_desc_sw128(smem_u32(sK + j * AGB_TILE));
    const uint64_t dd = umma_desc_sw128(smem_u32(sDO + i * AGB_TILE)), dv = umma_desc_sw128(smem_u32(sV + j * AGB_TILE));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem + AGB_C_S, umma_desc_advance_k(dq, k * 32), umma_desc_advance_k(dk, k * 32), idesc_kk, k != 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem + AGB_C_DP, umma_desc_advance_k(dd, k * 32), umma_desc_advance_k(dv, k * 32), idesc_kk, k != 0);
  };
  auto issue_grads = [&](int i, int j) {
    // dV_j += P^T dO_i ; dK_j += dS^T Q_i   (reduction over the 128 queries of tile i: 8 steps of 16 rows)
    const uint64_t ap = umma_desc_sw128_mn(smem_u32(sP), 128 * 128), as = umma_desc_sw128_mn(smem_u32(sDS), 128 * 128);
    const uint64_t bd = umma_desc_sw128(smem_u32(sDO + i * AGB_TILE)), bq = umma_desc_sw128(smem_u32(sQ + i * AGB_TILE));
#pragma unroll
    for (int k = 0; k < 8; ++k)
      umma_bf16(tmem + AGB_C_DV, umma_desc_advance_k(ap, k * 2048), umma_desc_advance_k(bd, k * 2048), idesc_ab, (i | k) != 0);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      umma_bf16(tmem + AGB_C_DK, umma_desc_advance_k(as, k * 2048), umma_desc_advance_k(bq, k * 2048), idesc_ab, (i | k) != 0);
    // dQ_i += dS K_j   (reduction over the 128 keys of tile j: 2 k-blocks of 64 keys x 4 steps)
    const uint64_t bk = umma_desc_sw128(smem_u32(sK + j * AGB_TILE));
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      const uint64_t a = umma_desc_sw128(smem_u32(sDS + kb * (128 * 128)));
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem + (i == 0 ? AGB_C_DQ0 : AGB_C_DQ1), umma_desc_advance_k(a, k * 32),
                  umma_desc_advance_k(bk, (kb * 4 + k) * 2048), idesc_b, (j | kb | k) != 0);
    }
  };

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, 4 * AGB_OPER);
    tma_load_2d(sQ, &tmQ, bar_load, h * ATT_HD, b * ATT_TP);
    tma_load_2d(sK, &tmK, bar_load, h * ATT_HD, b * ATT_TP);
    tma_load_2d(sV, &tmV, bar_load, v_col0 + h * ATT_HD, b * ATT_TP);
    tma_load_2d(sDO, &tmDO, bar_load, h * ATT_HD, b * ATT_TP);
    mbar_wait(bar_load, 0);
    tc_fence_after();
    issue_scores(0, 0);
    umma_commit(bar_mma);
  }
  __syncwarp();

  // per-row constants for both query tiles: D = rowsum(dO o O), lse
  // The four threads of a row (one per column quarter) each take 16 of the 64 head dims and exchange their partial sums
  // through shared memory (the P tile is not in use yet): a quarter of the global loads per thread (ncu: 21 % of the
  // kernel's stall samples were lg_throttle on these loads when every thread read the whole row).
  float Dv[2], Lv[2];
  {
    float* sD = reinterpret_cast<float*>(sP);   // [2 tiles][4 quarters][128 rows]
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const size_t grow = static_cast<size_t>(b) * ATT_TP + i * 128 + row;
      const uint4* po = reinterpret_cast<const uint4*>(O + grow * ldo + h * ATT_HD) + 2 * qtr;
      const uint4* pd = reinterpret_cast<const uint4*>(dO + grow * ldo + h * ATT_HD) + 2 * qtr;
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 a = __ldg(po + q), c = __ldg(pd + q);
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* hc = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 fa = __bfloat1622float2(ha[t]), fc = __bfloat1622float2(hc[t]);
          acc = fmaf(fa.x, fc.x, acc);
          acc = fmaf(fa.y, fc.y, acc);
        }
      }
      sD[(i * 4 + qtr) * 128 + row] = acc;
      Lv[i] = __ldg(lse + grow * ATT_HEADS + h) * 1.4426950408889634f;   // log2 domain
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i)
      Dv[i] = (sD[(i * 4 + 0) * 128 + row] + sD[(i * 4 + 1) * 128 + row]) + (sD[(i * 4 + 2) * 128 + row] + sD[(i * 4 + 3) * 128 + row]);
    __syncthreads();   // the first softmax overwrites the P tile
  }

  // inverse RoPE + bf16 store of a 64-column accumulator (this thread: row `r_in_win`, columns 16 qtr .. + 15)
  auto store_rot = [&](uint32_t tcol, int r_in_win, __nv_bfloat16* dst_base, int ld, int col0, bool rotate) {
    uint32_t r[16];
    tmem_ld_x16(tmem + t_row + tcol + qtr * 16, r);
    tmem_ld_wait();
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    const bool real = r_in_win < ATT_T;
    if (rotate && real) {
      const int which = r_in_win >> 7;   // r_in_win is row or 128 + row
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float4 cs = which ? pre_c[1][j] : pre_c[0][j], sn = which ? pre_s[1][j] : pre_s[0][j];
        const float cc[4] = {cs.x, cs.y, cs.z, cs.w}, ss[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float y1 = v[8 * j + 2 * t], y2 = v[8 * j + 2 * t + 1];
          v[8 * j + 2 * t] = y1 * cc[t] + y2 * ss[t];
          v[8 * j + 2 * t + 1] = -y1 * ss[t] + y2 * cc[t];
        }
      }
    }
    __nv_bfloat16* dst = dst_base + (static_cast<size_t>(b) * ATT_TP + r_in_win) * ld + col0 + h * ATT_HD + qtr * 16;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint4 o;
      o.x = real ? pack_bf16x2_att(v[8 * q], v[8 * q + 1]) : 0u;
      o.y = real ? pack_bf16x2_att(v[8 * q + 2], v[8 * q + 3]) : 0u;
      o.z = real ? pack_bf16x2_att(v[8 * q + 4], v[8 * q + 5]) : 0u;
      o.w = real ? pack_bf16x2_att(v[8 * q + 6], v[8 * q + 7]) : 0u;
      reinterpret_cast<uint4*>(dst)[q] = o;
    }
  };

  const float kscale = 0.125f * 1.4426950408889634f;
  // dropout of the attention weights (forward: attention.cuh): O = (P o m) V, so dV uses P o m and dP is masked too
  const uint32_t dthresh = drop ? drop->thresh : 0u;
  const float dinv = drop ? drop->inv_keep : 1.f;
  const uint32_t dkey = drop ? drop_key(drop->seed, drop_site) : 0u;
#pragma unroll 1
  for (int blk = 0; blk < 4; ++blk) {
    const int j = blk >> 1, i = blk & 1;
    mbar_wait(bar_mma, blk & 1);
    tc_fence_after();
    if (blk == 2) {   // key tile 0 is complete
      store_rot(AGB_C_DK, row, dKV, lddkv, 0, true);
      store_rot(AGB_C_DV, row, dKV, lddkv, v_col0, false);
    }
    const bool qreal = i * 128 + row < ATT_T;
    const float Dr = Dv[i], Lr = Lv[i];
    {
      const int col0 = qtr * 32;               // key column inside the tile
      uint32_t rs[32], rp[32];
      tmem_ld_x32(tmem + t_row + AGB_C_S + col0, rs);
      tmem_ld_x32(tmem + t_row + AGB_C_DP + col0, rp);
      tmem_ld_wait();
      uint32_t pp[16], ps[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        float p2[2], d2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int key = j * 128 + col0 + 2 * t + u;
          float p = exp2f(__uint_as_float(rs[2 * t + u]) * kscale - Lr);
          p = (key < ATT_T && qreal) ? p : 0.f;
          float mk = 1.f;
          if (dthresh != 0u)
            mk = drop_mul(dkey, ((static_cast<uint32_t>(b) * ATT_HEADS + h) * ATT_TP + i * 128 + row) * ATT_TP + key, dthresh, dinv);
          p2[u] = p * mk;
          d2[u] = p * (__uint_as_float(rp[2 * t + u]) * mk - Dr) * 0.125f;
        }
        pp[t] = pack_bf16x2_att(p2[0], p2[1]);
        ps[t] = pack_bf16x2_att(d2[0], d2[1]);
      }
      // [128 q][64 keys] k-block qtr / 2, 128B-swizzled rows; this thread's 32 columns = chunks 4 (qtr & 1) .. + 3
      uint8_t* bp = sP + (qtr >> 1) * (128 * 128);
      uint8_t* bs = sDS + (qtr >> 1) * (128 * 128);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        *reinterpret_cast<uint4*>(bp + sw128_offset(row, (qtr & 1) * 32 + 8 * q)) = make_uint4(pp[4 * q], pp[4 * q + 1], pp[4 * q + 2], pp[4 * q + 3]);
        *reinterpret_cast<uint4*>(bs + sw128_offset(row, (qtr & 1) * 32 + 8 * q)) = make_uint4(ps[4 * q], ps[4 * q + 1], ps[4 * q + 2], ps[4 * q + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      issue_grads(i, j);
      if (blk < 3) issue_scores((blk + 1) & 1, (blk + 1) >> 1);
      umma_commit(bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  store_rot(AGB_C_DK, 128 + row, dKV, lddkv, 0, true);
  store_rot(AGB_C_DV, 128 + row, dKV, lddkv, v_col0, false);
  store_rot(AGB_C_DQ0, row, dQ, lddq, 0, true);
  store_rot(AGB_C_DQ1, 128 + row, dQ, lddq, 0, true);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<AGB_TMEM_COLS>(tmem);
  }
}

// ------------------------------------------------------------------------------------------ local (tensor cores)
// LocalSelfAttention backward on tcgen05, one CTA per (window b, head h), built from attn_global_bwd_kernel's pieces.
// The 31 windows split by parity into two BLOCK-DIAGONAL problems: even windows w = 2m cover padded rows 16m .. 16m+15,
// odd windows w = 2m+1 cover 16m+8 .. 16m+23, i.e. aligned 16-row blocks once every operand is loaded 8 rows further
// down (TMA start coordinate; rows outside the 250 tokens are zero-filled by the TMA unit, which is exactly the
// reference's zero padding, model.py:425-428).  Inside a frame no window straddles a 128-row tile, so each of the four
// sub-problems (frame f in {0, 8}, tile i in {0, 1}) is five dense 128-wide products of the diagonal tile:
//     S = Q_i K_i^T, dP = dO_i V_i^T          (recomputed; only the 16 x 16 diagonal blocks are used)
//     P / dS                                  (CUDA cores: one 16-column softmax per row; everything else zero)
//     dV_i = P^T dO_i, dK_i = dS^T Q_i, dQ_i = dS K_i
// Results of the odd frame are added to the even frame's through global memory (bf16, same CTA, after a barrier).
// Output rows follow the reference's index shift: padded row p <-> token p - 3 for q / k / v, output row j = padded row j
// for dO (model.py:452-464); wn = 1 / (number of windows covering the output row), 0 for the dropped rows >= 250.
constexpr uint32_t ALT_C_S = 0, ALT_C_DP = 128, ALT_C_DQ = 256, ALT_C_DK = 320, ALT_C_DV = 384;

// tmQ / tmK / tmV / tmDO: 3-D maps {cols, 250 rows, B}, box {64, 128, 1} (make_tmap_3d).  The operands of one
// sub-problem (Q, K, V, dO of one 128-row tile: 64 KB) are double-buffered: the tile of sub-problem s + 1 is fetched while
// s is computed, and the global reads an epilogue needs (RoPE rows, the even frame's stored values) are issued before
// the wait on the products they follow.
constexpr int ALT_BUF = 4 * AGB_TILE;   // 64 KB

__global__ void __launch_bounds__(AGB_THREADS, 1)
attn_local_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, int v_col0,
                         const float* __restrict__ rope_cos, const float* __restrict__ rope_sin,
                         __nv_bfloat16* dQ, int lddq, __nv_bfloat16* dKV, int lddkv,
                         const DropParams* __restrict__ drop, uint32_t drop_site) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sOp = smem;                       // [2 buffers][Q | K | V | dO tiles of 128 rows]
  uint8_t* sP = sOp + 2 * ALT_BUF;
  uint8_t* sDS = sP + AGB_PS;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sDS + AGB_PS);   // [2]
  uint64_t* bar_mma = bar_load + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 3);

  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const uint32_t t_row = static_cast<uint32_t>(quad * 32) << 16;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init(&bar_load[0], 1);
    mbar_init(&bar_load[1], 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<AGB_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  constexpr uint32_t idesc_kk = umma_idesc_bf16(128, 128);
  constexpr uint32_t idesc_ab = umma_idesc_bf16_abmn(128, 64);
  constexpr uint32_t idesc_b = umma_idesc_bf16_bmn(128, 64);

  // sub-problem s: frame f = 8 (s >> 1), tile i = s & 1  ->  padded rows 128 i + f .. + 127
  auto load_sub = [&](int s) {   // elected thread
    const int r0 = (s & 1) * 128 + (s >> 1) * 8;
    uint8_t* buf = sOp + (s & 1) * ALT_BUF;
    uint64_t* bar = &bar_load[s & 1];
    mbar_arrive_expect_tx(bar, ALT_BUF);
    tma_load_3d(buf, &tmQ, bar, h * ATT_HD, r0 - 3, b);
    tma_load_3d(buf + AGB_TILE, &tmK, bar, h * ATT_HD, r0 - 3, b);
    tma_load_3d(buf + 2 * AGB_TILE, &tmV, bar, v_col0 + h * ATT_HD, r0 - 3, b);
    tma_load_3d(buf + 3 * AGB_TILE, &tmDO, bar, h * ATT_HD, r0, b);
  };
  auto issue_scores = [&](int s) {
    const uint8_t* buf = sOp + (s & 1) * ALT_BUF;
    const uint64_t dq = umma_desc_sw128(smem_u32(buf)), dk = umma_desc_sw128(smem_u32(buf + AGB_TILE));
    const uint64_t dv = umma_desc_sw128(smem_u32(buf + 2 * AGB_TILE)), dd = umma_desc_sw128(smem_u32(buf + 3 * AGB_TILE));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem + ALT_C_S, umma_desc_advance_k(dq, k * 32), umma_desc_advance_k(dk, k * 32), idesc_kk, k != 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem + ALT_C_DP, umma_desc_advance_k(dd, k * 32), umma_desc_advance_k(dv, k * 32), idesc_kk, k != 0);
  };
  auto issue_grads = [&](int s) {
    const uint8_t* buf = sOp + (s & 1) * ALT_BUF;
    const uint64_t ap = umma_desc_sw128_mn(smem_u32(sP), 128 * 128), as = umma_desc_sw128_mn(smem_u32(sDS), 128 * 128);
    const uint64_t bq = umma_desc_sw128(smem_u32(buf)), bk = umma_desc_sw128(smem_u32(buf + AGB_TILE));
    const uint64_t bd = umma_desc_sw128(smem_u32(buf + 3 * AGB_TILE));
#pragma unroll
    for (int k = 0; k < 8; ++k)
      umma_bf16(tmem + ALT_C_DV, umma_desc_advance_k(ap, k * 2048), umma_desc_advance_k(bd, k * 2048), idesc_ab, k != 0);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      umma_bf16(tmem + ALT_C_DK, umma_desc_advance_k(as, k * 2048), umma_desc_advance_k(bq, k * 2048), idesc_ab, k != 0);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      const uint64_t a = umma_desc_sw128(smem_u32(sDS + kb * (128 * 128)));
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem + ALT_C_DQ, umma_desc_advance_k(a, k * 32), umma_desc_advance_k(bk, (kb * 4 + k) * 2048), idesc_b, (kb | k) != 0);
    }
  };

  if (threadIdx.x == 0) {
    load_sub(0);
    load_sub(1);
  }
  // rows 250..255 of the gradient buffers are zero rows (they feed the padded dgrad / wgrad GEMMs)
  for (int idx = threadIdx.x; idx < (ATT_TP - ATT_T) * 32; idx += AGB_THREADS) {
    const int tok = ATT_T + (idx >> 5), cpair = idx & 31;
    const size_t g = static_cast<size_t>(b) * ATT_TP + tok;
    reinterpret_cast<uint32_t*>(dQ + g * lddq + h * ATT_HD)[cpair] = 0u;
    reinterpret_cast<uint32_t*>(dKV + g * lddkv + h * ATT_HD)[cpair] = 0u;
    reinterpret_cast<uint32_t*>(dKV + g * lddkv + v_col0 + h * ATT_HD)[cpair] = 0u;
  }
  if (threadIdx.x == 0) {
    mbar_wait(&bar_load[0], 0);
    tc_fence_after();
    issue_scores(0);
    umma_commit(bar_mma);
  }
  __syncwarp();

  // Epilogue of one sub-problem, in two parts.  fetch: the global reads (RoPE row of the token, and -- odd frame -- the
  // values the even frame stored), issued before the wait on the products.  emit: inverse RoPE, (+ stored), bf16 store of
  // this thread's 32 columns of dQ | dK | dV into token row p - 3.
  struct Pre {
    float4 cs[4], sn[4];
    uint4 old[3][4];
  };
  auto out_ptr = [&](int which, int tok) -> uint4* {
    __nv_bfloat16* base = which == 0 ? dQ : dKV;
    const int ld = which == 0 ? lddq : lddkv;
    const int col0 = which == 2 ? v_col0 : 0;
    return reinterpret_cast<uint4*>(base + (static_cast<size_t>(b) * ATT_TP + tok) * ld + col0 + h * ATT_HD + half * 32);
  };
  auto fetch = [&](int p, bool add, Pre& pre) {
    const int tok = p - 3;
    if (tok < 0 || tok >= ATT_T) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pre.cs[j] = __ldg(reinterpret_cast<const float4*>(rope_cos + tok * 32 + half * 16) + j);
      pre.sn[j] = __ldg(reinterpret_cast<const float4*>(rope_sin + tok * 32 + half * 16) + j);
    }
    if (add) {
#pragma unroll
      for (int wch = 0; wch < 3; ++wch) {
        const uint4* src = out_ptr(wch, tok);
#pragma unroll
        for (int q = 0; q < 4; ++q) pre.old[wch][q] = src[q];
      }
    }
  };
  auto emit = [&](int p, bool add, const Pre& pre) {
    const int tok = p - 3;
    const bool valid = tok >= 0 && tok < ATT_T;
#pragma unroll
    for (int wch = 0; wch < 3; ++wch) {
      uint32_t r[32];
      tmem_ld_x32(tmem + t_row + (wch == 0 ? ALT_C_DQ : wch == 1 ? ALT_C_DK : ALT_C_DV) + half * 32, r);
      tmem_ld_wait();
      if (valid) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (wch < 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float cc[4] = {pre.cs[j].x, pre.cs[j].y, pre.cs[j].z, pre.cs[j].w};
            const float ss[4] = {pre.sn[j].x, pre.sn[j].y, pre.sn[j].z, pre.sn[j].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float y1 = v[8 * j + 2 * t], y2 = v[8 * j + 2 * t + 1];
              v[8 * j + 2 * t] = y1 * cc[t] + y2 * ss[t];
              v[8 * j + 2 * t + 1] = -y1 * ss[t] + y2 * cc[t];
            }
          }
        }
        uint4* dst = out_ptr(wch, tok);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (add) {
            const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&pre.old[wch][q]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = __bfloat1622float2(ho[t]);
              v[8 * q + 2 * t] += f.x;
              v[8 * q + 2 * t + 1] += f.y;
            }
          }
          dst[q] = make_uint4(pack_bf16x2_att(v[8 * q], v[8 * q + 1]), pack_bf16x2_att(v[8 * q + 2], v[8 * q + 3]),
                              pack_bf16x2_att(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2_att(v[8 * q + 6], v[8 * q + 7]));
        }
      }
    }
  };

  const uint32_t dthresh = drop ? drop->thresh : 0u;
  const float dinv = drop ? drop->inv_keep : 1.f;
  const uint32_t dkey = drop ? drop_key(drop->seed, drop_site) : 0u;
  const bool active = (quad >> 1) == half;          // this row's 16-column block lies in this thread's 64-column half
  const int c0 = 4 * (quad & 1) + 2 * (lane >> 4);  // its first 8-column chunk inside the half
#pragma unroll 1
  for (int sub = 0; sub <= 4; ++sub) {
    // previous sub-problem's output row of this thread; its global reads go out before the wait
    const int pp = ((sub - 1) & 1) * 128 + ((sub - 1) >> 1) * 8 + row;
    const bool padd = sub >= 3;
    Pre pre;
    if (sub > 0) fetch(pp, padd, pre);
    mbar_wait(bar_mma, sub & 1);   // scores of this sub-problem and the gradients of the previous one
    tc_fence_after();
    if (threadIdx.x == 0 && sub >= 1 && sub + 1 < 4) load_sub(sub + 1);   // the buffer of sub - 1 is free now
    if (sub > 0) emit(pp, padd, pre);
    if (sub == 4) break;
    // ---- P and dS of this thread's row (padded row p): one window, 16 keys
    const int f = (sub >> 1) * 8, i = sub & 1;
    const int p = i * 128 + f + row;
    const int w = 2 * ((p - f) >> 4) + (f ? 1 : 0);
    uint32_t pk[8], sk[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { pk[t] = 0u; sk[t] = 0u; }
    if (active) {
      uint32_t rs[32], rp[32];
      tmem_ld_x32(tmem + t_row + ALT_C_S + quad * 32, rs);
      tmem_ld_x32(tmem + t_row + ALT_C_DP + quad * 32, rp);
      tmem_ld_wait();
      const bool hi = (lane & 16) != 0;
      float s[16], dp[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        s[k] = __uint_as_float(hi ? rs[16 + k] : rs[k]);
        dp[k] = __uint_as_float(hi ? rp[16 + k] : rp[k]);
      }
      const float wn = (p < ATT_T && w <= 30) ? ((p >= 8 && p < 248) ? 0.5f : 1.0f) : 0.f;
      float mx = s[0];
#pragma unroll
      for (int k = 1; k < 16; ++k) mx = fmaxf(mx, s[k]);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) { s[k] = __expf((s[k] - mx) * 0.125f); sum += s[k]; }
      const float inv = wn / sum;
      const uint32_t i0 = (((static_cast<uint32_t>(b) * ATT_HEADS + h) * 31u + static_cast<uint32_t>(w)) * 16u + static_cast<uint32_t>(p - 8 * w)) * 16u;
      float mk[16];
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        mk[k] = (dthresh != 0u) ? drop_mul(dkey, i0 + k, dthresh, dinv) : 1.f;
        s[k] *= inv;
        dp[k] *= mk[k];
        dot = fmaf(s[k], dp[k], dot);
      }
      const float mean = (wn > 0.f) ? dot / wn : 0.f;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        pk[t] = pack_bf16x2_att(s[2 * t] * mk[2 * t], s[2 * t + 1] * mk[2 * t + 1]);
        sk[t] = pack_bf16x2_att(s[2 * t] * (dp[2 * t] - mean) * 0.125f, s[2 * t + 1] * (dp[2 * t + 1] - mean) * 0.125f);
      }
    }
    {
      uint8_t* bp = sP + half * (128 * 128);
      uint8_t* bs = sDS + half * (128 * 128);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const bool a0 = active && q == c0, a1 = active && q == c0 + 1;
        uint4 vp = make_uint4(0u, 0u, 0u, 0u), vs = vp;
        if (a0) { vp = make_uint4(pk[0], pk[1], pk[2], pk[3]); vs = make_uint4(sk[0], sk[1], sk[2], sk[3]); }
        if (a1) { vp = make_uint4(pk[4], pk[5], pk[6], pk[7]); vs = make_uint4(sk[4], sk[5], sk[6], sk[7]); }
        *reinterpret_cast<uint4*>(bp + sw128_offset(row, 8 * q)) = vp;
        *reinterpret_cast<uint4*>(bs + sw128_offset(row, 8 * q)) = vs;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      issue_grads(sub);
      if (sub + 1 < 4) {
        mbar_wait(&bar_load[(sub + 1) & 1], static_cast<uint32_t>((sub + 1) >> 1));
        tc_fence_after();
        issue_scores(sub + 1);
      }
      umma_commit(bar_mma);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<AGB_TMEM_COLS>(tmem);
  }
}

}  // namespace a2m
