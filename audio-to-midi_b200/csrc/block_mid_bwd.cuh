// Backward of one narrow ConvNeXt Block (model.py:160-167) for C in {16, 32} with every product on tcgen05.
// block_small_bwd_kernel does the four per-token matrix-vector products (8 C^2 FMAs) and the two token-reduced outer
// products (4 C^2 FMAs) on the CUDA cores: 160 / 208 us per launch at 64 windows, a sixth of the training step.  Here the
// CUDA cores keep only what is per-token and narrow (depthwise k7, LayerNorm and its backward, GELU', the per-channel
// reductions, the transposed depthwise conv); thread t owns token tile0 - 3 + t == TMEM lane t, 122 inner tokens per tile.
//
//   1  coalesced load of x rows tile0-6 .. tile0+127 (fp32) into shared memory
//   2  thread t: xhat = LN(dwconv7(x)), a = xhat*lnw+lnb, dg2 = dOut*gamma  -> bf16 rows of the K-major tiles T0 (a), T1 (dg2)
//      and, masked to the inner tokens, of the wgrad operand TB = [m a | m dg2]
//   3  tcgen05  U [128 x 2C] = a W1^T        DH [128 x 2C] = dg2 W2            (recompute, and d gelu-output)
//   4  thread t: u + b1 -> gl = gelu(u), du = DH * gelu'(u)  -> bf16 rows of T0 (du), T1 (gl);  b1 gradient
//   5  tcgen05  DA [128 x C] = du W1         O [128 x C] = gl W2^T             (d LN-output, and pw2 output for d gamma)
//               WG [128 x 64] += [du | gl]^T [m a | m dg2]                       (both pointwise weight gradients in one
//               product over the tile's tokens: A and B MN-major straight from the row-per-token tiles; rows 0-63 x
//               columns 0..C-1 = dW1[h][c], rows 64-127 x columns C..2C-1 = dW2^T[h][c]; accumulated in TMEM over the tiles
//               of a persistent CTA and flushed once)
//   6  thread t: LayerNorm backward -> g = d(dwconv output); per-channel gradients by butterfly reductions
//   7  transposed depthwise conv of g (shared memory) + dOut -> dX for the inner tokens
//
// Parameter / gradient images: SmallBlockLayout<C> (fp32), as block_small_bwd_kernel.  wimg: bf16 tiles of 64-element
// 128B-swizzled rows  W1 [2C][c] | W2^T [2C][c] | W1^T [C][h] | W2 [C][h]  (a2m_api.cu pack_weights, train only).
#pragma once
#include "cnn_kernels.cuh"
#include "gemm_tc.cuh"
#include "gemm_wgrad.cuh"
#include "ptx.cuh"
#include "train_kernels.cuh"

namespace a2m {

// gelu_tanh and its derivative with the hardware tanh (one MUFU instead of ex2 + rcp; the forward's gelu_tanh_fast):
//   g = 0.5 x (1 + t),  dg = 0.5 (1 + t) + 0.5 x (1 - t^2) k (1 + 3 a x^2),  t = tanh(k (x + a x^3))
__device__ __forceinline__ void gelu_tanh_grad_fast(float x, float* g, float* dg) {
  const float k = 0.7978845608028654f, a = 0.044715f;
  const float x2 = x * x;
  const float u = k * x * fmaf(a, x2, 1.0f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hp = fmaf(0.5f, t, 0.5f);
  *g = x * hp;
  *dg = fmaf(0.5f * x * (k * fmaf(3.0f * a, x2, 1.0f)), fmaf(-t, t, 1.0f), hp);
}

constexpr int MBB_THREADS = 128;
constexpr int MBB_IN = MBB_THREADS - 6;

template <int C>
struct MidBwdCfg {
  static constexpr int H = 2 * C;
  static constexpr int RS = C + 4;
  static constexpr int TILE = 128 * 128;                          // one [128 tokens][64 bf16] swizzled tile
  static constexpr int W_BYTES = 6 * C * 128;                     // W1 (2C rows) | W2T (2C) | W1T (C) | W2 (C)
  static constexpr int X_BYTES = (MBB_THREADS + 6) * RS * 4;
  static constexpr int NP = 14 * C;                               // dw 7C | dwb | lnw | lnb | b1 2C | b2 | gamma
  static constexpr int NACC = 12 * C + H;                         // dw 7C | dwb | lnw | lnb | b2 | gamma | b1
  static constexpr int RAW = 1024 + 3 * TILE + ((W_BYTES + 1023) / 1024) * 1024 + X_BYTES + (NP + NACC) * 4 + 64;
  // DA / O re-use the columns of U / DH (drained by step 4 before step 5 is issued); WG lives across tiles.
  // C = 16: 128 TMEM columns, three CTAs per SM;  C = 32: 256 columns, two CTAs per SM (shared memory padded so that no
  // more CTAs than the tensor memory can serve become resident -- a CTA spinning in tcgen05.alloc would never finish).
  static constexpr int CTAS = C == 16 ? 3 : 2;
  static constexpr int MIN_SMEM = C == 16 ? 58 * 1024 : 80 * 1024;
  static constexpr size_t SMEM = RAW < MIN_SMEM ? MIN_SMEM : RAW;
  static constexpr uint32_t TMEM_COLS = C == 16 ? 128 : 256;
  static constexpr uint32_t COL_U = 0, COL_DH = H, COL_DA = 0, COL_O = C, COL_WG = 2 * H;
};

template <int C>
__global__ void __launch_bounds__(MBB_THREADS, MidBwdCfg<C>::CTAS)
block_mid_bwd_kernel(const float* Xin, const float* dOut, float* dX, int L, int M, const float* __restrict__ params,
                     const uint4* __restrict__ wimg, float* __restrict__ gparams) {
  using Cfg = MidBwdCfg<C>;
  using Lay = SmallBlockLayout<C>;
  static_assert(C == 16 || C == 32, "tensor-core narrow-stage backward: C in {16, 32}");
  constexpr int H = Cfg::H, RS = Cfg::RS, V = C / 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sT0 = smem;                                  // a, then du
  uint8_t* sT1 = sT0 + Cfg::TILE;                       // dg2, then gl
  uint8_t* sTB = sT1 + Cfg::TILE;                       // [m a | m dg2]
  uint8_t* sW = sTB + Cfg::TILE;
  uint8_t* sW1 = sW;                                    // [2C rows]  B of U  = a W1^T
  uint8_t* sW2T = sW1 + H * 128;                        // [2C rows]  B of DH = dg2 W2
  uint8_t* sW1T = sW2T + H * 128;                       // [C rows]   B of DA = du W1
  uint8_t* sW2 = sW1T + C * 128;                        // [C rows]   B of O  = gl W2^T
  float* sx = reinterpret_cast<float*>(sW + ((Cfg::W_BYTES + 1023) / 1024) * 1024);
  float* sp = sx + (MBB_THREADS + 6) * RS;
  float* sacc = sp + Cfg::NP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sacc + Cfg::NACC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  float* sg = reinterpret_cast<float*>(sT0);            // g rows (fp32, stride RS): aliases T0 / T1 once the products are done
  static_assert(MBB_THREADS * RS * 4 <= 2 * Cfg::TILE, "g rows fit in T0 | T1");
  // parameter image offsets inside sp
  constexpr int P_DW = 0, P_DWB = 7 * C, P_LNW = 8 * C, P_LNB = 9 * C, P_B1 = 10 * C, P_B2 = 12 * C, P_GAMMA = 13 * C;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = threadIdx.x;
  const uint32_t t_row = static_cast<uint32_t>(warp * 32) << 16;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  for (int i = threadIdx.x; i < 10 * C; i += MBB_THREADS) sp[i] = __ldg(params + Lay::DW + i);   // dw | dwb | lnw | lnb are contiguous
  for (int i = threadIdx.x; i < H; i += MBB_THREADS) sp[P_B1 + i] = __ldg(params + Lay::B1 + i);
  for (int i = threadIdx.x; i < C; i += MBB_THREADS) {
    sp[P_B2 + i] = __ldg(params + Lay::B2 + i);
    sp[P_GAMMA + i] = __ldg(params + Lay::GAMMA + i);
  }
  for (int i = threadIdx.x; i < Cfg::NACC; i += MBB_THREADS) sacc[i] = 0.f;
  copy_const_to_smem<Cfg::W_BYTES / 16, MBB_THREADS>(sW, wimg, threadIdx.x);
  // the operand tiles are only ever written in their first K (or 2C) columns: clear the rest once
  for (int i = threadIdx.x; i < 3 * Cfg::TILE / 16; i += MBB_THREADS) reinterpret_cast<uint4*>(sT0)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  constexpr uint32_t idesc_h = umma_idesc_bf16(128, H);           // U, DH
  constexpr uint32_t idesc_c = umma_idesc_bf16(128, C);           // DA, O
  constexpr uint32_t idesc_wg = umma_idesc_bf16_abmn(128, 64);    // WG

  const int ntiles = (M + MBB_IN - 1) / MBB_IN;
  uint32_t it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int tile0 = tile * MBB_IN;
    __syncthreads();   // the previous tile's readers of sx / sg are done
    // ---- 1: x rows tile0-6 .. tile0+127
    {
      constexpr int NV = (MBB_THREADS + 6) * V;
      constexpr int PER = (NV + MBB_THREADS - 1) / MBB_THREADS;
      float4 v[PER];
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * MBB_THREADS;
        const int r = i / V, q = i - r * V;
        const int g = tile0 - 6 + r;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < NV && g >= 0 && g < M) v[k] = reinterpret_cast<const float4*>(Xin + static_cast<size_t>(g) * C)[q];
      }
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * MBB_THREADS;
        const int r = i / V, q = i - r * V;
        if (i < NV) reinterpret_cast<float4*>(sx + r * RS)[q] = v[k];
      }
    }
    __syncthreads();
    const int tok = tile0 - 3 + row;                          // this thread's token (halo included)
    const bool inner = row >= 3 && row < MBB_IN + 3 && tok < M;
    const bool live = tok >= 0 && tok < M;
    const int tokc = min(max(tok, 0), M - 1);                 // dead threads compute on a clamped token, every contribution masked
    const float m = inner ? 1.f : 0.f;
    const int l = tokc % L;

    // ---- 2: dwconv7 + LN -> xhat (kept), a; dg2 = dOut * gamma
    float xhat[C];
    float inv;
    {
#pragma unroll
      for (int c = 0; c < C; ++c) xhat[c] = sp[P_DWB + c];
#pragma unroll
      for (int t = 0; t < 7; ++t) {
        const int ll = l + t - 3;
        if (ll >= 0 && ll < L) {
          const float* xr = sx + (row + t) * RS;              // sx row of token tok + t - 3
#pragma unroll
          for (int q = 0; q < V; ++q) {
            const float4 xv = reinterpret_cast<const float4*>(xr)[q];
            const float4 wv = reinterpret_cast<const float4*>(sp + P_DW + t * C)[q];
            xhat[4 * q] = fmaf(wv.x, xv.x, xhat[4 * q]);
            xhat[4 * q + 1] = fmaf(wv.y, xv.y, xhat[4 * q + 1]);
            xhat[4 * q + 2] = fmaf(wv.z, xv.z, xhat[4 * q + 2]);
            xhat[4 * q + 3] = fmaf(wv.w, xv.w, xhat[4 * q + 3]);
          }
        }
      }
      float mean = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) mean += xhat[c];
      mean *= (1.0f / C);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) { xhat[c] -= mean; var += xhat[c] * xhat[c]; }
      inv = rsqrtf(var * (1.0f / C) + kLnEps);
#pragma unroll
      for (int c = 0; c < C; ++c) xhat[c] *= inv;
      const float4* dsrc = reinterpret_cast<const float4*>(dOut + static_cast<size_t>(tokc) * C);
#pragma unroll
      for (int q8 = 0; q8 < C / 8; ++q8) {
        float a[8], d[8];
        const float4 d0 = dsrc[2 * q8], d1 = dsrc[2 * q8 + 1];
        d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = 8 * q8 + j;
          a[j] = xhat[c] * sp[P_LNW + c] + sp[P_LNB + c];
          d[j] *= sp[P_GAMMA + c];
        }
        *reinterpret_cast<uint4*>(sT0 + sw128_offset(row, 8 * q8)) =
            make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
        *reinterpret_cast<uint4*>(sT1 + sw128_offset(row, 8 * q8)) =
            make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
        *reinterpret_cast<uint4*>(sTB + sw128_offset(row, 8 * q8)) =
            make_uint4(pack_bf16x2(m * a[0], m * a[1]), pack_bf16x2(m * a[2], m * a[3]), pack_bf16x2(m * a[4], m * a[5]), pack_bf16x2(m * a[6], m * a[7]));
        *reinterpret_cast<uint4*>(sTB + sw128_offset(row, C + 8 * q8)) =
            make_uint4(pack_bf16x2(m * d[0], m * d[1]), pack_bf16x2(m * d[2], m * d[3]), pack_bf16x2(m * d[4], m * d[5]), pack_bf16x2(m * d[6], m * d[7]));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---- 3: U = a W1^T, DH = dg2 W2
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint64_t da = umma_desc_sw128(smem_u32(sT0)), dd = umma_desc_sw128(smem_u32(sT1));
      const uint64_t b1 = umma_desc_sw128(smem_u32(sW1)), b2 = umma_desc_sw128(smem_u32(sW2T));
#pragma unroll
      for (int k = 0; k < C / 16; ++k)
        umma_bf16(tmem + Cfg::COL_U, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(b1, k * 32), idesc_h, k != 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < C / 16; ++k)
        umma_bf16(tmem + Cfg::COL_DH, umma_desc_advance_k(dd, k * 32), umma_desc_advance_k(b2, k * 32), idesc_h, k != 0 ? 1u : 0u);
      umma_commit(&bars[0]);
    }
    __syncwarp();
    mbar_wait(&bars[0], it & 1);
    tc_fence_after();

    // ---- 4: gl = gelu(u), du = DH * gelu'(u) -> T0 (du), T1 (gl); b1 gradient (sum over the inner tokens of du)
    // C = 32: the fully unrolled kernel is > 100 KB of SASS and stalls on instruction fetch (ncu: stall_no_instruction ~ 1.1
    // warps per issue, profiles/r01f_block_mid_bwd_ncu.txt); the two large loop bodies are kept rolled there.
#pragma unroll(C == 32 ? 1 : 2)
    for (int c0 = 0; c0 < H; c0 += 32) {
      uint32_t ru[32], rd[32];
      tmem_ld_x32(tmem + t_row + Cfg::COL_U + c0, ru);
      tmem_ld_x32(tmem + t_row + Cfg::COL_DH + c0, rd);
      tmem_ld_wait();
      float du[32];
      uint32_t pg[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float g0, g1, d0, d1;
        gelu_tanh_grad_fast(__uint_as_float(ru[2 * j]) + sp[P_B1 + c0 + 2 * j], &g0, &d0);
        gelu_tanh_grad_fast(__uint_as_float(ru[2 * j + 1]) + sp[P_B1 + c0 + 2 * j + 1], &g1, &d1);
        du[2 * j] = __uint_as_float(rd[2 * j]) * d0;
        du[2 * j + 1] = __uint_as_float(rd[2 * j + 1]) * d1;
        pg[j] = pack_bf16x2(g0, g1);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        *reinterpret_cast<uint4*>(sT0 + sw128_offset(row, c0 + 8 * q)) =
            make_uint4(pack_bf16x2(du[8 * q], du[8 * q + 1]), pack_bf16x2(du[8 * q + 2], du[8 * q + 3]),
                       pack_bf16x2(du[8 * q + 4], du[8 * q + 5]), pack_bf16x2(du[8 * q + 6], du[8 * q + 7]));
        *reinterpret_cast<uint4*>(sT1 + sw128_offset(row, c0 + 8 * q)) = make_uint4(pg[4 * q], pg[4 * q + 1], pg[4 * q + 2], pg[4 * q + 3]);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) du[j] *= m;
      const float r = warp_vec_reduce<32>(du, lane);
      atomicAdd(&sacc[12 * C + c0 + lane], r);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---- 5: DA = du W1, O = gl W2^T, WG += [du | gl]^T [m a | m dg2]
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint64_t du_k = umma_desc_sw128(smem_u32(sT0)), gl_k = umma_desc_sw128(smem_u32(sT1));
      const uint64_t b3 = umma_desc_sw128(smem_u32(sW1T)), b4 = umma_desc_sw128(smem_u32(sW2));
#pragma unroll
      for (int k = 0; k < H / 16; ++k)
        umma_bf16(tmem + Cfg::COL_DA, umma_desc_advance_k(du_k, k * 32), umma_desc_advance_k(b3, k * 32), idesc_c, k != 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < H / 16; ++k)
        umma_bf16(tmem + Cfg::COL_O, umma_desc_advance_k(gl_k, k * 32), umma_desc_advance_k(b4, k * 32), idesc_c, k != 0 ? 1u : 0u);
      const uint64_t wa = umma_desc_sw128_mn(smem_u32(sT0), Cfg::TILE), wb = umma_desc_sw128(smem_u32(sTB));
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16(tmem + Cfg::COL_WG, umma_desc_advance_k(wa, k * 2048), umma_desc_advance_k(wb, k * 2048), idesc_wg, (it | k) != 0 ? 1u : 0u);
      umma_commit(&bars[1]);
    }
    __syncwarp();
    mbar_wait(&bars[1], it & 1);
    tc_fence_after();

    // ---- 6: LayerNorm backward -> g; per-channel gradients
    float g[C];
    {
      float da[C], o[C];
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 16) {
        uint32_t r1[16], r2[16];
        tmem_ld_x16(tmem + t_row + Cfg::COL_DA + c0, r1);
        tmem_ld_x16(tmem + t_row + Cfg::COL_O + c0, r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          da[c0 + j] = __uint_as_float(r1[j]);
          o[c0 + j] = __uint_as_float(r2[j]) + sp[P_B2 + c0 + j];
        }
      }
      const int ch = vec_reduce_channel<C>(lane);
      const bool lead = (lane % (32 / C)) == 0;
      {
        float v1[C], v2[C], v3[C], v4[C];
        const float4* dsrc = reinterpret_cast<const float4*>(dOut + static_cast<size_t>(tokc) * C);
#pragma unroll
        for (int q = 0; q < V; ++q) {
          const float4 dv = dsrc[q];
          const float d[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = 4 * q + j;
            v1[c] = m * da[c] * xhat[c];
            v2[c] = m * da[c];
            v3[c] = m * d[j] * sp[P_GAMMA + c];
            v4[c] = m * d[j] * o[c];
          }
        }
        const float r1 = warp_vec_reduce<C>(v1, lane), r2 = warp_vec_reduce<C>(v2, lane);
        const float r3 = warp_vec_reduce<C>(v3, lane), r4 = warp_vec_reduce<C>(v4, lane);
        if (lead) {
          atomicAdd(&sacc[8 * C + ch], r1);
          atomicAdd(&sacc[9 * C + ch], r2);
          atomicAdd(&sacc[10 * C + ch], r3);
          atomicAdd(&sacc[11 * C + ch], r4);
        }
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        da[c] *= sp[P_LNW + c];
        s1 += da[c];
        s2 += da[c] * xhat[c];
      }
      s1 *= (1.0f / C);
      s2 *= (1.0f / C);
#pragma unroll
      for (int c = 0; c < C; ++c) g[c] = live ? inv * (da[c] - s1 - xhat[c] * s2) : 0.f;
      // depthwise-conv gradients: dw[t][c] += g[c] x[tok + t - 3][c], dwb[c] += g[c]
#pragma unroll(C == 32 ? 1 : 7)
      for (int t = 0; t < 7; ++t) {
        const int ll = l + t - 3;
        const float mt = (ll >= 0 && ll < L) ? m : 0.f;
        const float* xr = sx + (row + t) * RS;
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = mt * g[c] * xr[c];
        const float r = warp_vec_reduce<C>(v, lane);
        if (lead) atomicAdd(&sacc[t * C + ch], r);
      }
      float v[C];
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = m * g[c];
      const float r = warp_vec_reduce<C>(v, lane);
      if (lead) atomicAdd(&sacc[7 * C + ch], r);
    }
    // every product has completed (bars[1]): T0 / T1 are free, g rows go there
#pragma unroll
    for (int q = 0; q < V; ++q) reinterpret_cast<float4*>(sg + row * RS)[q] = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
    __syncthreads();
    // ---- 7: dX = dOut + transposed depthwise conv of g, inner tokens
    if (inner) {
      float acc[C];
      const float4* dsrc = reinterpret_cast<const float4*>(dOut + static_cast<size_t>(tok) * C);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        const float4 v = dsrc[q];
        acc[4 * q] = v.x; acc[4 * q + 1] = v.y; acc[4 * q + 2] = v.z; acc[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int t = 0; t < 7; ++t) {
        const int ll = l - t + 3;
        if (ll >= 0 && ll < L) {
          const float* gr = sg + (row - t + 3) * RS;
#pragma unroll
          for (int q = 0; q < V; ++q) {
            const float4 gv = reinterpret_cast<const float4*>(gr)[q];
            const float4 wv = reinterpret_cast<const float4*>(sp + P_DW + t * C)[q];
            acc[4 * q] = fmaf(wv.x, gv.x, acc[4 * q]);
            acc[4 * q + 1] = fmaf(wv.y, gv.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(wv.z, gv.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(wv.w, gv.w, acc[4 * q + 3]);
          }
        }
      }
      float4* dst = reinterpret_cast<float4*>(dX + static_cast<size_t>(tok) * C);
#pragma unroll
      for (int q = 0; q < V; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
    // sg dirties bytes of T0 / T1; every column a product of the next tile reads as K is rewritten first (steps 2, 4),
    // and the MN columns beyond 2C that WG reads (C = 16) only reach accumulator rows that are never flushed.
  }

  // ---- flush: pointwise weight gradients out of TMEM (lane = h or 64 + h), per-channel gradients out of sacc
  if (it > 0) {
    const int hrow = row & 63;
    const bool first = row < 64;
    uint32_t r[32];
    if constexpr (C == 32) {
      tmem_ld_x32(tmem + t_row + Cfg::COL_WG + (first ? 0 : 32), r);
    } else {
      tmem_ld_x32(tmem + t_row + Cfg::COL_WG, r);
    }
    tmem_ld_wait();
    if (hrow < H) {
      float* dst = gparams + (first ? Lay::W1 : Lay::W2T) + hrow * C;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int col = (C == 32) ? c : (first ? c : 16 + c);
        atomicAdd(dst + c, __uint_as_float(r[col]));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cfg::NACC; i += MBB_THREADS) {
    // sacc order: dw[7][C] | dwb | lnw | lnb | b2 | gamma | b1[H]  ->  SmallBlockLayout offsets
    const int seg = i / C, c = i % C;
    const int off = seg < 7 ? Lay::DW + i : seg == 7 ? Lay::DWB + c : seg == 8 ? Lay::LNW + c : seg == 9 ? Lay::LNB + c
                  : seg == 10 ? Lay::B2 + c : seg == 11 ? Lay::GAMMA + c : Lay::B1 + (i - 12 * C);
    atomicAdd(gparams + off, sacc[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem);
  }
}

// ------------------------------------------------------------------------------------------ narrow Downsample backward
// Backward of Downsample (model.py:102-118) for CIN in {16, 32} on tcgen05.  Forward (down_mid_kernel): n = [LN(x_2t) | LN(x_2t+1)]
// (K2 = 2 CIN values), y = W n + b.  downsample_small_bwd_kernel spends 2 K2 COUT FMAs per output token on the CUDA cores
// (85 / 155 us per launch).  Here thread t owns output token tile0 + t == TMEM lane t (no halo):
//   1  x pair and dY row straight from global memory (each thread reads two contiguous 4 K2-byte / 4 COUT-byte runs)
//   2  xhat, inv of both tokens (kept); n -> bf16 row of TN; dY -> bf16 row of TDY; validity -> column 0 of TONE
//   3  tcgen05  DN [128 x K2] = dY W                                   (A = TDY K-major, B = W^T rows [k][o])
//               WG [128 x 80] += dY^T [n | 1]                          (A = TDY MN-major, B = TN | TONE MN-major, N = 64 + 16:
//               rows 0..COUT-1 x columns 0..K2-1 = dW[o][k], column 64 = db[o]; accumulated in TMEM over the CTA's tiles)
//   4  thread t: LayerNorm backward of both tokens out of TMEM -> dX rows; dlnw / dlnb accumulated in registers
// wimg: bf16 W^T as K2 rows of 64 swizzled elements ([k][o], a2m_api.cu pack_weights, train only).
template <int CIN>
struct DownMidBwdCfg {
  static constexpr int COUT = 2 * CIN, K2 = 2 * CIN;
  static constexpr int TILE = 128 * 128;
  static constexpr int W_BYTES = K2 * 128;
  static constexpr int RAW = 1024 + 3 * TILE + W_BYTES + 2 * CIN * 4 + 64;
  static constexpr size_t SMEM = RAW < 80 * 1024 ? 80 * 1024 : RAW;   // at most two CTAs per SM (256 TMEM columns each)
  static constexpr uint32_t TMEM_COLS = 256;
  static constexpr uint32_t COL_DN = 0, COL_WG = 64;    // WG: 80 columns
};

template <int CIN>
__global__ void __launch_bounds__(MBB_THREADS, 2)
down_mid_bwd_kernel(const float* X, const float* dY, float* dX, int M_out, const float* __restrict__ params,
                    const uint4* __restrict__ wimg, float* __restrict__ gparams) {
  using Cfg = DownMidBwdCfg<CIN>;
  using Lay = SmallDownLayout<CIN>;
  static_assert(CIN == 16 || CIN == 32, "tensor-core narrow Downsample backward: CIN in {16, 32}");
  constexpr int COUT = Cfg::COUT, K2 = Cfg::K2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sDY = smem;                       // [128 tok][COUT]   A of DN (K-major) and of WG (MN-major)
  uint8_t* sN = sDY + Cfg::TILE;             // [128 tok][K2]     B of WG, columns 0..63
  uint8_t* sOne = sN + Cfg::TILE;            // [128 tok][col 0 = valid]   B of WG, columns 64..79
  uint8_t* sWT = sOne + Cfg::TILE;           // [K2 rows][COUT]   B of DN
  float* sln = reinterpret_cast<float*>(sWT + Cfg::W_BYTES);   // lnw | lnb
  uint64_t* bars = reinterpret_cast<uint64_t*>(sln + 2 * CIN);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = threadIdx.x;
  const uint32_t t_row = static_cast<uint32_t>(warp * 32) << 16;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  for (int i = threadIdx.x; i < 2 * CIN; i += MBB_THREADS) sln[i] = __ldg(params + Lay::LNW + i);   // lnw | lnb are contiguous
  copy_const_to_smem<Cfg::W_BYTES / 16, MBB_THREADS>(sWT, wimg, threadIdx.x);
  for (int i = threadIdx.x; i < 3 * Cfg::TILE / 16; i += MBB_THREADS) reinterpret_cast<uint4*>(sDY)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  constexpr uint32_t idesc_dn = umma_idesc_bf16(128, K2);
  constexpr uint32_t idesc_wg = umma_idesc_bf16_abmn(128, 80);
  float glw[CIN], glb[CIN];
#pragma unroll
  for (int c = 0; c < CIN; ++c) { glw[c] = 0.f; glb[c] = 0.f; }

  const int ntiles = (M_out + MBB_THREADS - 1) / MBB_THREADS;
  uint32_t it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int tok = tile * MBB_THREADS + row;
    const bool valid = tok < M_out;
    const int tokc = min(tok, M_out - 1);
    // ---- 1, 2
    float xh[K2], inv[2];
    {
      const float4* src = reinterpret_cast<const float4*>(X + static_cast<size_t>(tokc) * K2);
#pragma unroll
      for (int q = 0; q < K2 / 4; ++q) {
        const float4 v = src[q];
        xh[4 * q] = v.x; xh[4 * q + 1] = v.y; xh[4 * q + 2] = v.z; xh[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float mean = 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) mean += xh[t * CIN + c];
        mean *= (1.0f / CIN);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) { xh[t * CIN + c] -= mean; var += xh[t * CIN + c] * xh[t * CIN + c]; }
        inv[t] = rsqrtf(var * (1.0f / CIN) + kLnEps);
#pragma unroll
        for (int c = 0; c < CIN; ++c) xh[t * CIN + c] *= inv[t];
      }
      const float mv = valid ? 1.f : 0.f;
#pragma unroll
      for (int q = 0; q < K2 / 8; ++q) {
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = (8 * q + j) % CIN;
          a[j] = mv * (xh[8 * q + j] * sln[c] + sln[CIN + c]);
        }
        *reinterpret_cast<uint4*>(sN + sw128_offset(row, 8 * q)) =
            make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
      }
      const float4* dsrc = reinterpret_cast<const float4*>(dY + static_cast<size_t>(tokc) * COUT);
#pragma unroll
      for (int q = 0; q < COUT / 8; ++q) {
        const float4 d0 = dsrc[2 * q], d1 = dsrc[2 * q + 1];
        *reinterpret_cast<uint4*>(sDY + sw128_offset(row, 8 * q)) =
            make_uint4(pack_bf16x2(mv * d0.x, mv * d0.y), pack_bf16x2(mv * d0.z, mv * d0.w), pack_bf16x2(mv * d1.x, mv * d1.y),
                       pack_bf16x2(mv * d1.z, mv * d1.w));
      }
      *reinterpret_cast<uint4*>(sOne + sw128_offset(row, 0)) = make_uint4(valid ? 0x00003f80u : 0u, 0u, 0u, 0u);   // bf16 1.0 in column 0
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- 3
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint64_t da = umma_desc_sw128(smem_u32(sDY)), db = umma_desc_sw128(smem_u32(sWT));
#pragma unroll
      for (int k = 0; k < COUT / 16; ++k)
        umma_bf16(tmem + Cfg::COL_DN, umma_desc_advance_k(da, k * 32), umma_desc_advance_k(db, k * 32), idesc_dn, k != 0 ? 1u : 0u);
      // A: one 64-wide MN block (dY; the instruction's second block of 64 M rows reads TN: those accumulator rows are never flushed)
      const uint64_t wa = umma_desc_sw128_mn(smem_u32(sDY), Cfg::TILE), wb = umma_desc_sw128_mn(smem_u32(sN), Cfg::TILE);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16(tmem + Cfg::COL_WG, umma_desc_advance_k(wa, k * 2048), umma_desc_advance_k(wb, k * 2048), idesc_wg, (it | k) != 0 ? 1u : 0u);
      umma_commit(&bars[0]);
    }
    __syncwarp();
    mbar_wait(&bars[0], it & 1);
    tc_fence_after();
    // ---- 4: LayerNorm backward of the two input tokens
    {
      float* dst = dX + static_cast<size_t>(tokc) * K2;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float dn[CIN];
#pragma unroll
        for (int c0 = 0; c0 < CIN; c0 += 16) {
          uint32_t r[16];
          tmem_ld_x16(tmem + t_row + Cfg::COL_DN + t * CIN + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) dn[c0 + j] = __uint_as_float(r[j]);
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float d = dn[c];                 // zero for invalid tokens (their dY row was zeroed)
          glw[c] += d * xh[t * CIN + c];
          glb[c] += d;
          dn[c] = d * sln[c];
          s1 += dn[c];
          s2 += dn[c] * xh[t * CIN + c];
        }
        s1 *= (1.0f / CIN);
        s2 *= (1.0f / CIN);
        if (valid) {
#pragma unroll
          for (int q = 0; q < CIN / 4; ++q) {
            float4 o;
            o.x = inv[t] * (dn[4 * q] - s1 - xh[t * CIN + 4 * q] * s2);
            o.y = inv[t] * (dn[4 * q + 1] - s1 - xh[t * CIN + 4 * q + 1] * s2);
            o.z = inv[t] * (dn[4 * q + 2] - s1 - xh[t * CIN + 4 * q + 2] * s2);
            o.w = inv[t] * (dn[4 * q + 3] - s1 - xh[t * CIN + 4 * q + 3] * s2);
            reinterpret_cast<float4*>(dst + t * CIN)[q] = o;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // every thread has drained DN and the products have read the tiles: the next tile may overwrite both
  }

  // ---- flush
  if (it > 0) {
    // WG lane o (< COUT): columns 0..K2-1 = dW[o][k], column 64 = db[o]
    uint32_t r[32];
#pragma unroll
    for (int c0 = 0; c0 < K2; c0 += 32) {
      tmem_ld_x32(tmem + t_row + Cfg::COL_WG + c0, r);
      tmem_ld_wait();
      if (row < COUT) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(gparams + Lay::W + row * K2 + c0 + j, __uint_as_float(r[j]));
      }
    }
    uint32_t rb[16];
    tmem_ld_x16(tmem + t_row + Cfg::COL_WG + 64, rb);
    tmem_ld_wait();
    if (row < COUT) atomicAdd(gparams + Lay::B + row, __uint_as_float(rb[0]));
  }
  {
    const float r1 = warp_vec_reduce<CIN>(glw, lane), r2 = warp_vec_reduce<CIN>(glb, lane);
    if ((lane % (32 / CIN)) == 0) {
      atomicAdd(gparams + Lay::LNW + vec_reduce_channel<CIN>(lane), r1);
      atomicAdd(gparams + Lay::LNB + vec_reduce_channel<CIN>(lane), r2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem);
  }
}

}  // namespace a2m
