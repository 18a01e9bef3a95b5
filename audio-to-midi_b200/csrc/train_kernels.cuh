// CUDA-core kernels of the training path (train.py:39-62, 259-332): loss gradient, the memory-bound backward glue
// (LayerNorm, GELU / GLU, depthwise conv), the small-channel CNN stages, gradient (un)packing and AdamW.
// Parameter gradients are accumulated in fp32 into a "packed gradient" buffer that mirrors the packed weight
// images the forward kernels read (same layouts), and are scattered back to the reference's leaf layout by
// grad_unpack_kernel.  Per-channel reductions keep per-thread register partials over a grid-stride loop and
// finish with one shared-memory reduction and one atomicAdd per channel per CTA.
#pragma once
#include "cnn_kernels.cuh"
#include "gemm_tc.cuh"

namespace a2m {

// d/dx of gelu_tanh(x) = x * sigmoid(2u), u = k (x + a x^3)
__device__ __forceinline__ void gelu_tanh_grad(float x, float* g, float* dg) {
  const float k = 0.7978845608028654f, a = 0.044715f;
  const float u = k * (x + a * x * x * x);
  const float s = __fdividef(1.0f, 1.0f + __expf(-2.0f * u));
  *g = x * s;
  *dg = s + x * s * (1.0f - s) * 2.0f * k * (1.0f + 3.0f * a * x * x);
}

__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// ------------------------------------------------------------------------------------------ loss
// logits, labels [B, 250, 90] fp32 -> dz [B*256, 128] bf16 (zero in the padding) = (sigmoid(z) - y) * gscale,
// loss[0] += sum BCEWithLogits(z, y) * lscale   (optax.sigmoid_binary_cross_entropy, train.py:43-47, 61-62)
__global__ void __launch_bounds__(256) bce_grad_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                       __nv_bfloat16* __restrict__ dz, float* __restrict__ loss, int B,
                                                       float gscale, float lscale) {
  const int idx = blockIdx.x * 256 + threadIdx.x;   // over B*256*128
  float l = 0.f;
  if (idx < B * 256 * 128) {
    const int col = idx & 127, row = (idx >> 7) & 255, b = idx >> 15;
    float d = 0.f;
    if (col < 90 && row < 250) {
      const size_t o = (static_cast<size_t>(b) * 250 + row) * 90 + col;
      const float z = logits[o], y = labels[o];
      const float p = __fdividef(1.0f, 1.0f + __expf(-z));
      d = (p - y) * gscale;
      l = (fmaxf(z, 0.f) - z * y + log1pf(__expf(-fabsf(z)))) * lscale;
    }
    dz[idx] = op1_rn(d);
  }
  l = warp_sum(l);
  __shared__ float sl[8];
  if ((threadIdx.x & 31) == 0) sl[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sl[i];
    atomicAdd(loss, t);
  }
}

// Arbitrary cotangent of the logits (the custom_vjp backward of a JAX caller): dlogits [B, 250, 90] fp32 -> dz [B*256, 128]
// bf16, zero in the padding.
__global__ void __launch_bounds__(256) dlogits_pack_kernel(const float* __restrict__ dlogits, __nv_bfloat16* __restrict__ dz, int B) {
  const int idx = blockIdx.x * 256 + threadIdx.x;   // over B*256*128
  if (idx >= B * 256 * 128) return;
  const int col = idx & 127, row = (idx >> 7) & 255, b = idx >> 15;
  float d = 0.f;
  if (col < 90 && row < 250) d = dlogits[(static_cast<size_t>(b) * 250 + row) * 90 + col];
  dz[idx] = op1_rn(d);
}

// Per-window validation loss (train.py:99-102 testset_loss_function): out[b] = sum_{t,c} BCEWithLogits(z, y), one CTA per window.
__global__ void __launch_bounds__(256) bce_window_loss_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                              float* __restrict__ out, int per_window) {
  const size_t base = static_cast<size_t>(blockIdx.x) * per_window;
  float l = 0.f;
  for (int i = threadIdx.x; i < per_window; i += 256) {
    const float z = logits[base + i], y = labels[base + i];
    l += fmaxf(z, 0.f) - z * y + log1pf(__expf(-fabsf(z)));
  }
  l = warp_sum(l);
  __shared__ float sl[8];
  if ((threadIdx.x & 31) == 0) sl[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sl[i];
    out[blockIdx.x] = t;
  }
}

// ------------------------------------------------------------------------------------------ column sums (bias grads)
// out[n] += sum_rows dY[row, n]   (bf16 [M, N], N % 8 == 0, N <= 1024)
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ dY, int ld, int M, int N,
                                                          float* __restrict__ out) {
  const int groups = N / 8;                 // uint4 column groups
  const int rows_per_it = 256 / groups > 0 ? 256 / groups : 1;
  // thread -> (column group cg, row lane rl); N <= 1024 -> groups <= 128
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (rl < rows_per_it) {
    for (int r = blockIdx.x * rows_per_it + rl; r < M; r += gridDim.x * rows_per_it) {
      const uint4 v = *reinterpret_cast<const uint4*>(dY + static_cast<size_t>(r) * ld + cg * 8);
      acc[0] += bf16lo(v.x); acc[1] += bf16hi(v.x); acc[2] += bf16lo(v.y); acc[3] += bf16hi(v.y);
      acc[4] += bf16lo(v.z); acc[5] += bf16hi(v.z); acc[6] += bf16lo(v.w); acc[7] += bf16hi(v.w);
    }
  }
  __shared__ float s[256 * 8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += 256) {
    float t = 0.f;
    const int g = c / 8, j = c % 8;
    for (int r = 0; r < rows_per_it; ++r) t += s[(r * groups + g) * 8 + j];
    atomicAdd(out + c, t);
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm backward (rows)
// y = xhat * w + b over a row of C channels (model.py:100,117,162,190,539,546,759).  One warp per row, grid-stride.
//   X    fp32 rows (compact [B*Lx] layout), dY fp32 rows ([B*Ly] layout; row t of window b is valid for t < Lx)
//   dX   fp32 [B*Lx]: ACCUMULATE ? dX += dx : dX = dx;  dX16 optional bf16 copy of the final dX
//   gw / gb: fp32 [C] accumulated with atomics
template <int C, bool ACCUMULATE>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* X, const float* dY, int nB, int Lx, int Ly,
                                                     const float* __restrict__ lnw, float* dX, __nv_bfloat16* dX16,
                                                     float* __restrict__ gw, float* __restrict__ gb,
                                                     const DropParams* __restrict__ drop, uint32_t drop_site) {
  using RM = RowMap<C>;
  constexpr int PER = RM::PER;
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float lw[PER], aw[PER], ab[PER];
  RM::load(lnw, lane, lw);
#pragma unroll
  for (int j = 0; j < PER; ++j) { aw[j] = 0.f; ab[j] = 0.f; }
  pdl_wait();
  const int rows = nB * Lx;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const int b = r / Lx, t = r - b * Lx;
    float x[PER], dy[PER];
    RM::load(X + static_cast<size_t>(r) * C, lane, x);
    RM::load(dY + (static_cast<size_t>(b) * Ly + t) * C, lane, dy);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) s += x[j];
    const float mean = warp_sum(s) * (1.0f / C);
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) { x[j] -= mean; v += x[j] * x[j]; }
    const float inv = rsqrtf(warp_sum(v) * (1.0f / C) + kLnEps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      x[j] *= inv;                       // xhat
      aw[j] += dy[j] * x[j];
      ab[j] += dy[j];
      dy[j] *= lw[j];                    // d xhat
      s1 += dy[j];
      s2 += dy[j] * x[j];
    }
    s1 = warp_sum(s1) * (1.0f / C);
    s2 = warp_sum(s2) * (1.0f / C);
    float dx[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) dx[j] = inv * (dy[j] - s1 - x[j] * s2);
    float* dst = dX + static_cast<size_t>(r) * C;
    if constexpr (ACCUMULATE) {
      float old[PER];
      RM::load(dst, lane, old);
#pragma unroll
      for (int j = 0; j < PER; ++j) dx[j] += old[j];
    }
    RM::store_f32(dst, lane, dx);
    if (dX16 != nullptr) {
      // the bf16 copy feeds the FFN of the layer below, whose output went through dropout (model.py:237):
      // d(pre-dropout) = mask / (1 - p) * d(post-dropout); the fp32 copy is the residual path and stays unmasked
      if (drop != nullptr && drop->thresh != 0u) {
        const uint32_t key = drop_key(drop->seed, drop_site), th = drop->thresh;
        const float inv_keep = drop->inv_keep;
#pragma unroll
        for (int g = 0; g < RM::G; ++g)
#pragma unroll
          for (int j = 0; j < RM::VW; ++j)
            dx[g * RM::VW + j] *= drop_mul(key, static_cast<uint32_t>(r) * C + RM::chan(lane, g) + j, th, inv_keep);
      }
      RM::store_bf16(dX16 + static_cast<size_t>(r) * C, lane, dx);
    }
  }
  __shared__ float sw[8][C], sb[8][C];
#pragma unroll
  for (int g = 0; g < RM::G; ++g)
#pragma unroll
    for (int j = 0; j < RM::VW; ++j) {
      sw[warp][RM::chan(lane, g) + j] = aw[g * RM::VW + j];
      sb[warp][RM::chan(lane, g) + j] = ab[g * RM::VW + j];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float tw = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { tw += sw[w][c]; tb += sb[w][c]; }
    atomicAdd(gw + c, tw);
    atomicAdd(gb + c, tb);
  }
}

// ------------------------------------------------------------------------------------------ GELU / GLU
// h = gelu(u)  (bf16, n8 = elements / 8)
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const __nv_bfloat16* U, __nv_bfloat16* Hh, size_t n8) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n8; i += gridDim.x * 256ull) {
    const uint4 u = reinterpret_cast<const uint4*>(U)[i];
    uint4 o;
    o.x = pack_bf16x2(gelu_tanh_f(bf16lo(u.x)), gelu_tanh_f(bf16hi(u.x)));
    o.y = pack_bf16x2(gelu_tanh_f(bf16lo(u.y)), gelu_tanh_f(bf16hi(u.y)));
    o.z = pack_bf16x2(gelu_tanh_f(bf16lo(u.z)), gelu_tanh_f(bf16hi(u.z)));
    o.w = pack_bf16x2(gelu_tanh_f(bf16lo(u.w)), gelu_tanh_f(bf16hi(u.w)));
    reinterpret_cast<uint4*>(Hh)[i] = o;
  }
}
// du = dh * gelu'(u)   (in place over dh allowed)
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat16* U, const __nv_bfloat16* dH, __nv_bfloat16* dU, size_t n8) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n8; i += gridDim.x * 256ull) {
    const uint4 u = reinterpret_cast<const uint4*>(U)[i];
    const uint4 d = reinterpret_cast<const uint4*>(dH)[i];
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w}, dd[4] = {d.x, d.y, d.z, d.w};
    uint32_t oo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float g0, g1, d0, d1;
      gelu_tanh_grad(bf16lo(uu[q]), &g0, &d0);
      gelu_tanh_grad(bf16hi(uu[q]), &g1, &d1);
      oo[q] = pack_bf16x2(bf16lo(dd[q]) * d0, bf16hi(dd[q]) * d1);
    }
    reinterpret_cast<uint4*>(dU)[i] = make_uint4(oo[0], oo[1], oo[2], oo[3]);
  }
}
// FeedForwardBlock gate (model.py:233-234) in the tile-permuted layout of the packed FFN-1 weight: U [rows, 2F] holds,
// per 256-column tile tb, 128 "gelu" columns then their 128 "gate" columns; H [rows, F] column tb*128 + r.
__global__ void __launch_bounds__(256) glu_fwd_kernel(const __nv_bfloat16* U, __nv_bfloat16* Hh, int rows, int F) {
  pdl_launch_dependents();
  pdl_wait();
  const int per_row = F / 8;
  const size_t n = static_cast<size_t>(rows) * per_row;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) {
    const int r = static_cast<int>(i / per_row), c8 = static_cast<int>(i % per_row) * 8;
    const int tb = c8 >> 7, rr = c8 & 127;
    const __nv_bfloat16* up = U + static_cast<size_t>(r) * 2 * F + tb * 256 + rr;
    const uint4 a = *reinterpret_cast<const uint4*>(up), g = *reinterpret_cast<const uint4*>(up + 128);
    uint4 o;
    o.x = pack_bf16x2(gelu_tanh_f(bf16lo(a.x)) * bf16lo(g.x), gelu_tanh_f(bf16hi(a.x)) * bf16hi(g.x));
    o.y = pack_bf16x2(gelu_tanh_f(bf16lo(a.y)) * bf16lo(g.y), gelu_tanh_f(bf16hi(a.y)) * bf16hi(g.y));
    o.z = pack_bf16x2(gelu_tanh_f(bf16lo(a.z)) * bf16lo(g.z), gelu_tanh_f(bf16hi(a.z)) * bf16hi(g.z));
    o.w = pack_bf16x2(gelu_tanh_f(bf16lo(a.w)) * bf16lo(g.w), gelu_tanh_f(bf16hi(a.w)) * bf16hi(g.w));
    *reinterpret_cast<uint4*>(Hh + static_cast<size_t>(r) * F + c8) = o;
  }
}
// dU (same permuted layout as U) from dH [rows, F]
__global__ void __launch_bounds__(256) glu_bwd_kernel(const __nv_bfloat16* U, const __nv_bfloat16* dH, __nv_bfloat16* dU, int rows, int F) {
  pdl_launch_dependents();
  pdl_wait();
  const int per_row = F / 8;
  const size_t n = static_cast<size_t>(rows) * per_row;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) {
    const int r = static_cast<int>(i / per_row), c8 = static_cast<int>(i % per_row) * 8;
    const int tb = c8 >> 7, rr = c8 & 127;
    const size_t uo = static_cast<size_t>(r) * 2 * F + tb * 256 + rr;
    const uint4 a = *reinterpret_cast<const uint4*>(U + uo), g = *reinterpret_cast<const uint4*>(U + uo + 128);
    const uint4 d = *reinterpret_cast<const uint4*>(dH + static_cast<size_t>(r) * F + c8);
    const uint32_t aa[4] = {a.x, a.y, a.z, a.w}, gg[4] = {g.x, g.y, g.z, g.w}, dd[4] = {d.x, d.y, d.z, d.w};
    uint32_t o1[4], o2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float g0, g1, d0, d1;
      gelu_tanh_grad(bf16lo(aa[q]), &g0, &d0);
      gelu_tanh_grad(bf16hi(aa[q]), &g1, &d1);
      const float dl = bf16lo(dd[q]), dh = bf16hi(dd[q]);
      o1[q] = pack_bf16x2(dl * bf16lo(gg[q]) * d0, dh * bf16hi(gg[q]) * d1);
      o2[q] = pack_bf16x2(dl * g0, dh * g1);
    }
    *reinterpret_cast<uint4*>(dU + uo) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    *reinterpret_cast<uint4*>(dU + uo + 128) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  }
}

__global__ void set_drop_params_kernel(DropParams* p, uint32_t seed, uint32_t thresh, float inv_keep) {
  p->seed = seed; p->thresh = thresh; p->inv_keep = inv_keep; p->pad = 0u;
}

// fp32 -> bf16 copy (n4 = elements / 4)
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* X, __nv_bfloat16* Y, size_t n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n4; i += gridDim.x * 256ull) {
    const float4 v = reinterpret_cast<const float4*>(X)[i];
    reinterpret_cast<uint2*>(Y)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ------------------------------------------------------------------------------------------ Block (C >= 64): dwconv + LN backward
// Forward (model.py:160-162): y = dwconv7(x) + b ; a = LN(y).  Given dA (grad wrt a, fp32) and dOut (grad wrt the Block
// output == grad wrt the residual branch input), produces
//     dX = dOut + dwconv7^T( LN'(dA) )      (fp32 + bf16 copy)
// and accumulates the parameter gradients in the packed image layout  dw[7][C] | dwb[C] | lnw[C] | lnb[C].
// Tile of DWB_TOK tokens per iteration; g = LN'(dA) is recomputed for 3 halo rows on each side.
// Each tile's inputs (x rows with a 6-row halo, dA rows with a 3-row halo, dOut rows) are contiguous byte ranges of global
// memory: one elected thread fetches them with three 1-D bulk async copies (cp.async.bulk, mbarrier completion) into a
// two-stage shared-memory ring, one tile ahead of the warps that consume them, so the loop never waits on a dependent
// global load (the first version did three times per tile: 80 % of its stall samples, profiles/r01d_dwconv_bwd.txt).
template <int C>
struct DwBwd {
  static constexpr int THREADS = (C == 256) ? 256 : 512;     // 16 warps hide the shuffle / smem latency of the row reductions
  static constexpr int NW = THREADS / 32;
  static constexpr int TOK = 4096 / C;                      // 64 / 32 / 16 tokens per tile: ~58 KB per stage
  static constexpr int XR = TOK + 12, AR = TOK + 6;         // rows of x / dA per tile
  static constexpr int STAGE_FLOATS = (XR + AR + TOK) * C;
  static constexpr size_t SMEM = static_cast<size_t>(2 * STAGE_FLOATS + AR * C) * 4 + 64;   // + g rows + 2 mbarriers
  static_assert(NW * 10 * C * 4 <= 2 * STAGE_FLOATS * 4, "the final reduction re-uses the stage buffers");
};
template <int C>
constexpr size_t dwconv_ln_bwd_smem() { return DwBwd<C>::SMEM; }

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int C>
__global__ void __launch_bounds__(DwBwd<C>::THREADS) dwconv_ln_bwd_kernel(const float* X, const float* dA, const float* dOut, float* dX,
                                                                    __nv_bfloat16* dX16, int L, int M,
                                                                    const float* __restrict__ params, float* __restrict__ gparams) {
  using RM = RowMap<C>;
  using D = DwBwd<C>;
  constexpr int PER = RM::PER, TOK = D::TOK, XR = D::XR, AR = D::AR, DWB_THREADS = D::THREADS;
  extern __shared__ __align__(128) float smem_dwb[];   // own name: the other kernels of this unit declare smem_f with 16-byte alignment
  float* stage0 = smem_dwb;
  float* sg = smem_dwb + 2 * D::STAGE_FLOATS;              // g rows tile0-3 .. tile0+TOK+2
  uint64_t* bars = reinterpret_cast<uint64_t*>(sg + AR * C);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  float w[7][PER], bias[PER], lw[PER];
#pragma unroll
  for (int t = 0; t < 7; ++t) RM::load(params + t * C, lane, w[t]);
  RM::load(params + 7 * C, lane, bias);
  RM::load(params + 8 * C, lane, lw);
  float gdw[7][PER], gdb[PER], glw[PER], glb[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    gdb[j] = glw[j] = glb[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 7; ++t) gdw[t][j] = 0.f;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  const int ntiles = (M + TOK - 1) / TOK;
  // issues the three bulk copies of `tile` into stage `s`; rows outside [0, M) are not copied (zero-filled by the consumer)
  auto issue = [&](int tile, int s) {
    const int tile0 = tile * TOK;
    float* sx = stage0 + s * D::STAGE_FLOATS;
    float* sda = sx + XR * C;
    float* sdo = sda + AR * C;
    const int x0 = max(tile0 - 6, 0), x1 = min(tile0 + TOK + 6, M);
    const int a0 = max(tile0 - 3, 0), a1 = min(tile0 + TOK + 3, M);
    const int o0 = tile0, o1 = min(tile0 + TOK, M);
    const uint32_t bytes = static_cast<uint32_t>((x1 - x0) + (a1 - a0) + (o1 - o0)) * C * 4u;
    mbar_arrive_expect_tx(&bars[s], bytes);
    bulk_load_1d(sx + (x0 - (tile0 - 6)) * C, X + static_cast<size_t>(x0) * C, static_cast<uint32_t>(x1 - x0) * C * 4u, &bars[s]);
    bulk_load_1d(sda + (a0 - (tile0 - 3)) * C, dA + static_cast<size_t>(a0) * C, static_cast<uint32_t>(a1 - a0) * C * 4u, &bars[s]);
    bulk_load_1d(sdo, dOut + static_cast<size_t>(o0) * C, static_cast<uint32_t>(o1 - o0) * C * 4u, &bars[s]);
  };
  if (threadIdx.x == 0 && static_cast<int>(blockIdx.x) < ntiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const int tile0 = tile * TOK;
    float* sx = stage0 + s * D::STAGE_FLOATS;
    float* sda = sx + XR * C;
    float* sdo = sda + AR * C;
    // prefetch the next tile into the other stage: its previous contents were consumed before the barrier that ended
    // the previous iteration
    if (threadIdx.x == 0 && tile + static_cast<int>(gridDim.x) < ntiles) issue(tile + gridDim.x, s ^ 1);
    mbar_wait(&bars[s], (it >> 1) & 1);
    if (tile0 < 6 || tile0 + TOK + 6 > M) {   // edge tiles: rows outside the tensor were not copied
      for (int i = threadIdx.x; i < XR * (C / 4); i += DWB_THREADS) {
        const int r = i / (C / 4), g = tile0 - 6 + r;
        if (g < 0 || g >= M) reinterpret_cast<float4*>(sx + r * C)[i % (C / 4)] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
    }
    // phase A: g rows tile0-3 .. tile0+TOK+2
    // position inside the window of this warp's rows: one runtime modulo per tile and phase, then l += NW (mod L) per row
    int la = (tile0 - 3 + warp + L) % L;
    for (int lr = warp; lr < AR; lr += DWB_THREADS / 32, la += DWB_THREADS / 32, la -= (la >= L) ? L : 0) {
      const int tok = tile0 - 3 + lr;
      float g[PER];
#pragma unroll
      for (int j = 0; j < PER; ++j) g[j] = 0.f;
      if (tok >= 0 && tok < M) {
        const int l = la;
        float y[PER], xr[7][PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) y[j] = bias[j];
#pragma unroll
        for (int t = 0; t < 7; ++t) {
          const int ll = l + t - 3;
          if (ll >= 0 && ll < L) {
            RM::load(sx + (lr + t) * C, lane, xr[t]);   // sx row of token tok + t - 3 is (lr + 3) + (t - 3)
#pragma unroll
            for (int j = 0; j < PER; ++j) y[j] = fmaf(w[t][j], xr[t][j], y[j]);
          } else {
#pragma unroll
            for (int j = 0; j < PER; ++j) xr[t][j] = 0.f;
          }
        }
        // single-pass statistics (two independent shuffle reductions instead of two dependent ones)
        float sm = 0.f, sq = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) { sm += y[j]; sq = fmaf(y[j], y[j], sq); }
        sm = warp_sum(sm);
        sq = warp_sum(sq);
        const float mean = sm * (1.0f / C);
        const float inv = rsqrtf(fmaxf(sq * (1.0f / C) - mean * mean, 0.f) + kLnEps);
#pragma unroll
        for (int j = 0; j < PER; ++j) y[j] -= mean;
        float da[PER];
        RM::load(sda + lr * C, lane, da);
        const bool inner = lr >= 3 && lr < TOK + 3;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          y[j] *= inv;
          if (inner) { glw[j] += da[j] * y[j]; glb[j] += da[j]; }
          da[j] *= lw[j];
          s1 += da[j];
          s2 += da[j] * y[j];
        }
        s1 = warp_sum(s1) * (1.0f / C);
        s2 = warp_sum(s2) * (1.0f / C);
#pragma unroll
        for (int j = 0; j < PER; ++j) g[j] = inv * (da[j] - s1 - y[j] * s2);
        if (inner) {
#pragma unroll
          for (int j = 0; j < PER; ++j) {
            gdb[j] += g[j];
#pragma unroll
            for (int t = 0; t < 7; ++t) gdw[t][j] = fmaf(g[j], xr[t][j], gdw[t][j]);
          }
        }
      }
      RM::store_f32(sg + lr * C, lane, g);
    }
    __syncthreads();
    // phase B: dX[m] = dOut[m] + sum_t w[t] g[m - t + 3]  (conv positions inside the same window only)
    int lb = (tile0 + warp) % L;
    for (int lr = warp; lr < TOK; lr += DWB_THREADS / 32, lb += DWB_THREADS / 32, lb -= (lb >= L) ? L : 0) {
      const int tok = tile0 + lr;
      if (tok >= M) break;
      const int l = lb;
      float acc[PER];
      RM::load(sdo + lr * C, lane, acc);
#pragma unroll
      for (int t = 0; t < 7; ++t) {
        const int ll = l - t + 3;
        if (ll >= 0 && ll < L) {
          float gr[PER];
          RM::load(sg + (lr + 3 - t + 3) * C, lane, gr);
#pragma unroll
          for (int j = 0; j < PER; ++j) acc[j] = fmaf(w[t][j], gr[j], acc[j]);
        }
      }
      RM::store_f32(dX + static_cast<size_t>(tok) * C, lane, acc);
      if (dX16 != nullptr) RM::store_bf16(dX16 + static_cast<size_t>(tok) * C, lane, acc);
    }
    __syncthreads();   // stage s and sg are free again
  }
  // parameter gradients: registers -> smem [warp][10][C] (re-using the stage buffers) -> atomics
  float* sred = stage0;
  float* mine = sred + warp * 10 * C;
#pragma unroll
  for (int g = 0; g < RM::G; ++g)
#pragma unroll
    for (int j = 0; j < RM::VW; ++j) {
      const int c = RM::chan(lane, g) + j, jj = g * RM::VW + j;
#pragma unroll
      for (int t = 0; t < 7; ++t) mine[t * C + c] = gdw[t][jj];
      mine[7 * C + c] = gdb[jj];
      mine[8 * C + c] = glw[jj];
      mine[9 * C + c] = glb[jj];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 10 * C; i += DWB_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < D::NW; ++wv) t += sred[wv * 10 * C + i];
    atomicAdd(gparams + i, t);
  }
}

// Finishes the layer-scale chain rule of a Block (model.py:166-167: out = x + gamma * (W2 h + b2)) after the tensor-core
// wgrad produced, with the UNSCALED block-output gradient as dY,  G[c,k] = sum_t dOut[t,c] h[t,k]  (in the dW2 slot)
// and g[c] = sum_t dOut[t,c] (in the db2 slot):   dgamma[c] = sum_k W2[c,k] G[c,k] + b2[c] g[c];  dW2 = gamma G;  db2 = gamma g.
__global__ void __launch_bounds__(256) block_gamma_finish_kernel(const __nv_bfloat16* __restrict__ W2, const float* __restrict__ b2,
                                                                 const float* __restrict__ gamma, float* G, float* gb2, float* ggamma,
                                                                 int C, int H) {
  const int c = blockIdx.x;
  const float gm = gamma[c];
  float acc = 0.f;
  for (int k = threadIdx.x; k < H; k += 256) {
    const float gv = G[static_cast<size_t>(c) * H + k];
    acc = fmaf(__bfloat162float(W2[static_cast<size_t>(c) * H + k]), gv, acc);
    G[static_cast<size_t>(c) * H + k] = gm * gv;
  }
  acc = warp_sum(acc);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s[i];
    const float g = gb2[c];
    ggamma[c] += t + b2[c] * g;
    gb2[c] = gm * g;
  }
}

// ------------------------------------------------------------------------------------------ small Block backward (C <= 32)
// One thread per token recomputes the Block (block_small_kernel) and its backward; tile = SBB_IN inner tokens + 3 halo
// tokens on each side whose g = d(dwconv output) is needed by the transposed depthwise conv.  Parameter gradients are
// written in the SmallBlockLayout image.
//   * pointwise-conv weight gradients: per-token du / gelu(u) / a / dg2 are staged row-per-token in shared memory (bf16,
//     like every tensor-core wgrad operand) and reduced over the tile's tokens by register-tiled outer products
//     (4 x 4 outputs per thread, 2 vector loads per 16 FMAs);
//   * per-channel gradients (dw taps, biases, LN affine, gamma): butterfly vector reduction over the warp
//     (1 shuffle per value instead of 5), then one shared-memory atomic per channel per warp.
constexpr int SBB_THREADS = 128;
constexpr int SBB_IN = SBB_THREADS - 6;

// Sums v[c] over the 32 lanes for every c < C with C-1 (+ log2(32/C)) shuffles.  Returns, in lane l, the total of channel
// vec_reduce_channel<C>(l); for C < 32 every group of 32/C consecutive lanes holds the same channel.
template <int C>
__device__ __forceinline__ int vec_reduce_channel(int lane) { return lane / (32 / C); }
template <int C>
__device__ __forceinline__ float warp_vec_reduce(float (&v)[C], int lane) {
  int off = 16;
#pragma unroll
  for (int n = C; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? v[i] : v[i + n / 2];
      const float keep = upper ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (int o = (32 / C) / 2; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
  return v[0];
}

template <int C>
struct SmallBwdSmem {
  using Lay = SmallBlockLayout<C>;
  static constexpr int H = 2 * C;
  static constexpr int RS = (C == 4) ? 4 : C + 4;
  static constexpr int P = (Lay::TOTAL + 3) & ~3;
  static constexpr int SX = (SBB_THREADS + 6) * RS;        // rows tile0-6 .. tile0+IN+5  (IN + 12 = THREADS + 6)
  static constexpr int SG = SBB_THREADS * RS;              // g rows tile0-3 .. tile0+IN+2
  static constexpr int HS = H + 4;                         // bf16 row strides (8-byte aligned, bank-staggered)
  static constexpr int CS = C + 4;
  static constexpr int SV16 = SBB_THREADS * (2 * HS + 2 * CS);   // du | gl | a | dg2, row per token, bf16
  static constexpr size_t BYTES = static_cast<size_t>(P + SX + SG) * 4 + static_cast<size_t>(SV16) * 2;
};

// acc[TH][TC] += sum_tok A[tok][r0 .. r0+TH) x B[tok][c0 .. c0+TC)   (bf16 rows in shared memory)
template <int TH, int TC, int AS, int BS>
__device__ __forceinline__ void outer_tile(const __nv_bfloat16* sA, const __nv_bfloat16* sB, int r0, int c0, float (&acc)[TH][TC]) {
#pragma unroll 4
  for (int t = 0; t < SBB_THREADS; ++t) {
    float a[TH], b[TC];
    if constexpr (TH == 4) {
      const uint2 v = *reinterpret_cast<const uint2*>(sA + t * AS + r0);
      a[0] = bf16lo(v.x); a[1] = bf16hi(v.x); a[2] = bf16lo(v.y); a[3] = bf16hi(v.y);
    } else if constexpr (TH == 2) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(sA + t * AS + r0);
      a[0] = bf16lo(v); a[1] = bf16hi(v);
    } else {
#pragma unroll
      for (int i = 0; i < TH; ++i) a[i] = __bfloat162float(sA[t * AS + r0 + i]);
    }
    if constexpr (TC == 4) {
      const uint2 v = *reinterpret_cast<const uint2*>(sB + t * BS + c0);
      b[0] = bf16lo(v.x); b[1] = bf16hi(v.x); b[2] = bf16lo(v.y); b[3] = bf16hi(v.y);
    } else if constexpr (TC == 2) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(sB + t * BS + c0);
      b[0] = bf16lo(v); b[1] = bf16hi(v);
    } else {
#pragma unroll
      for (int i = 0; i < TC; ++i) b[i] = __bfloat162float(sB[t * BS + c0 + i]);
    }
#pragma unroll
    for (int i = 0; i < TH; ++i)
#pragma unroll
      for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// C <= 8: capped at 128 registers for four CTAs per SM (C = 4: 104 -> 83 us; C = 8 needs 105 and must not be given more --
// with a minimum of one CTA the compiler took more registers and the kernel went from 102 to 143 us; five CTAs at C = 8 and
// a larger grid for the narrow Downsample backward were measured neutral / slower)
template <int C>
__global__ void __launch_bounds__(SBB_THREADS, C <= 8 ? 4 : 0) block_small_bwd_kernel(const float* Xin, const float* dOut, float* dX, int L, int M,
                                                                      const float* __restrict__ params, float* __restrict__ gparams) {
  using Lay = SmallBlockLayout<C>;
  using SM = SmallBwdSmem<C>;
  constexpr int H = Lay::H, RS = SM::RS, V = C / 4, HS = SM::HS, CS = SM::CS;
  constexpr int TH = (C >= 8) ? C / 8 : 1, TC = TH;   // outer-product register tile: 4x4 (C = 32), 2x2, 1x1
  constexpr int NTILE = (H / TH) * (C / TC);          // 128 threads own a tile (32 for C = 4)
  extern __shared__ __align__(16) float smem_f[];
  float* sp = smem_f;
  float* sx = sp + SM::P;
  float* sg = sx + SM::SX;
  __nv_bfloat16* sdu = reinterpret_cast<__nv_bfloat16*>(sg + SM::SG);   // [tok][HS]
  __nv_bfloat16* sgl = sdu + SBB_THREADS * HS;
  __nv_bfloat16* sa = sgl + SBB_THREADS * HS;                           // [tok][CS]
  __nv_bfloat16* sdg = sa + SBB_THREADS * CS;
  // CTA accumulator of the per-channel gradients  [dw 7C | dwb | lnw | lnb | b2 | gamma | b1 (H)]
  __shared__ float sacc[12 * C + H];
  for (int i = threadIdx.x; i < 12 * C + H; i += SBB_THREADS) sacc[i] = 0.f;
  const int lane = threadIdx.x & 31;
  const int my_r0 = (threadIdx.x / (C / TC)) * TH, my_c0 = (threadIdx.x % (C / TC)) * TC;
  float accW1[TH][TC], accW2[TH][TC];
#pragma unroll
  for (int i = 0; i < TH; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) { accW1[i][j] = 0.f; accW2[i][j] = 0.f; }
  // C = 4: the two 8 x 4 pointwise weight gradients are accumulated per thread in registers over all of its tokens
  // (64 FMAs per token) and reduced over the CTA once, instead of one warp walking the staged rows of every tile
  // (~2000 instructions per tile on the critical path of the other three warps).
  constexpr bool REG_WG = (C == 4);
  constexpr int RW = REG_WG ? H * C : 1;
  float regW1[RW], regW2[RW];
#pragma unroll
  for (int i = 0; i < RW; ++i) { regW1[i] = 0.f; regW2[i] = 0.f; }
  // C <= 8: the twelve per-channel sums (dw 7 taps, dwb, lnw, lnb, b2, gamma) stay in a register per lane over the CTA's
  // tiles and reach shared memory once, instead of twelve shared-memory atomics per warp and tile (ncu: 14 % of the stall
  // samples of block_small_bwd_kernel<4>, profiles/r01f_train_stalls_by_source_line_2.txt)
  constexpr bool REG_CH = (C <= 8);
  float rch[REG_CH ? 12 : 1];
#pragma unroll
  for (int i = 0; i < (REG_CH ? 12 : 1); ++i) rch[i] = 0.f;

  for (int i = threadIdx.x; i < Lay::TOTAL; i += SBB_THREADS) sp[i] = __ldg(params + i);
  const int ntiles = (M + SBB_IN - 1) / SBB_IN;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tile0 = tile * SBB_IN;
    __syncthreads();
    for (int i = threadIdx.x; i < (SBB_THREADS + 6) * V; i += SBB_THREADS) {
      const int r = i / V, q = i - r * V;
      const int g = tile0 - 6 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g >= 0 && g < M) v = reinterpret_cast<const float4*>(Xin + static_cast<size_t>(g) * C)[q];
      reinterpret_cast<float4*>(sx + r * RS)[q] = v;
    }
    __syncthreads();
    const int tok = tile0 - 3 + threadIdx.x;                 // this thread's token (halo included)
    const bool inner = threadIdx.x >= 3 && threadIdx.x < SBB_IN + 3 && tok < M;
    const bool live = tok >= 0 && tok < M;
    const int tokc = min(max(tok, 0), M - 1);   // dead threads compute on a clamped token with every contribution masked
    const float m = inner ? 1.f : 0.f;
    const int l = tokc % L;
    float g[C];
    {
      float y[C];
#pragma unroll
      for (int c = 0; c < C; ++c) y[c] = sp[Lay::DWB + c];
#pragma unroll
      for (int t = 0; t < 7; ++t) {
        const int ll = l + t - 3;
        if (ll >= 0 && ll < L) {
          const float* row = sx + (threadIdx.x + t) * RS;    // sx row of token tok + t - 3
#pragma unroll
          for (int c = 0; c < C; ++c) y[c] = fmaf(sp[Lay::DW + t * C + c], row[c], y[c]);
        }
      }
      float mean = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) mean += y[c];
      mean *= (1.0f / C);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) { y[c] -= mean; var += y[c] * y[c]; }
      const float inv = rsqrtf(var * (1.0f / C) + kLnEps);
      float a[C], dout[C], dg2[C], da[C], o[C];
      const float4* dsrc = reinterpret_cast<const float4*>(dOut + static_cast<size_t>(tokc) * C);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        const float4 v = dsrc[q];
        dout[4 * q] = v.x; dout[4 * q + 1] = v.y; dout[4 * q + 2] = v.z; dout[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        y[c] *= inv;                                          // xhat
        a[c] = y[c] * sp[Lay::LNW + c] + sp[Lay::LNB + c];
        dg2[c] = dout[c] * sp[Lay::GAMMA + c];
        da[c] = 0.f;
        o[c] = sp[Lay::B2 + c];
      }
      if constexpr (!REG_WG) {
#pragma unroll
        for (int q = 0; q < V; ++q) {
          *reinterpret_cast<uint2*>(sa + threadIdx.x * CS + 4 * q) =
              make_uint2(pack_bf16x2(m * a[4 * q], m * a[4 * q + 1]), pack_bf16x2(m * a[4 * q + 2], m * a[4 * q + 3]));
          *reinterpret_cast<uint2*>(sdg + threadIdx.x * CS + 4 * q) =
              make_uint2(pack_bf16x2(m * dg2[4 * q], m * dg2[4 * q + 1]), pack_bf16x2(m * dg2[4 * q + 2], m * dg2[4 * q + 3]));
        }
      }
      float b1part = 0.f;   // this lane's share of the b1 gradient is reduced below, 4 hidden units at a time
#pragma unroll(REG_WG ? 2 : 1)
      for (int h0 = 0; h0 < H; h0 += 4) {
        float du4[4], gl4[4];
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          const int h = h0 + hh;
          float u = sp[Lay::B1 + h], dh = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            u = fmaf(sp[Lay::W1 + h * C + c], a[c], u);
            dh = fmaf(sp[Lay::W2T + h * C + c], dg2[c], dh);
          }
          float gl, dgl;
          gelu_tanh_grad(u, &gl, &dgl);
          const float du = dh * dgl;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            da[c] = fmaf(sp[Lay::W1 + h * C + c], du, da[c]);
            o[c] = fmaf(sp[Lay::W2T + h * C + c], gl, o[c]);
          }
          du4[hh] = m * du;
          gl4[hh] = m * gl;
        }
        if constexpr (REG_WG) {
#pragma unroll
          for (int hh = 0; hh < 4; ++hh)
#pragma unroll
            for (int c = 0; c < C; ++c) {
              regW1[(h0 + hh) * C + c] = fmaf(du4[hh], a[c], regW1[(h0 + hh) * C + c]);
              regW2[(h0 + hh) * C + c] = fmaf(gl4[hh], dg2[c], regW2[(h0 + hh) * C + c]);
            }
        } else {
          *reinterpret_cast<uint2*>(sdu + threadIdx.x * HS + h0) = make_uint2(pack_bf16x2(du4[0], du4[1]), pack_bf16x2(du4[2], du4[3]));
          *reinterpret_cast<uint2*>(sgl + threadIdx.x * HS + h0) = make_uint2(pack_bf16x2(gl4[0], gl4[1]), pack_bf16x2(gl4[2], gl4[3]));
        }
        // b1 gradient: sum over the warp's tokens of du for these 4 hidden units
        const float r = warp_vec_reduce<4>(du4, lane);
        if ((lane & 7) == 0) atomicAdd(&sacc[12 * C + h0 + (lane >> 3)], r);
        (void)b1part;
      }
      // per-channel sums over tokens: lnw, lnb, b2, gamma
      {
        float v1[C], v2[C], v3[C], v4[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          v1[c] = m * da[c] * y[c];
          v2[c] = m * da[c];
          v3[c] = m * dg2[c];
          v4[c] = m * dout[c] * o[c];
        }
        const int ch = vec_reduce_channel<C>(lane);
        const bool lead = (lane % (32 / C)) == 0;
        const float r1 = warp_vec_reduce<C>(v1, lane), r2 = warp_vec_reduce<C>(v2, lane);
        const float r3 = warp_vec_reduce<C>(v3, lane), r4 = warp_vec_reduce<C>(v4, lane);
        if constexpr (REG_CH) {
          rch[8] += r1; rch[9] += r2; rch[10] += r3; rch[11] += r4;
        } else if (lead) {
          atomicAdd(&sacc[8 * C + ch], r1);
          atomicAdd(&sacc[9 * C + ch], r2);
          atomicAdd(&sacc[10 * C + ch], r3);
          atomicAdd(&sacc[11 * C + ch], r4);
        }
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        da[c] *= sp[Lay::LNW + c];
        s1 += da[c];
        s2 += da[c] * y[c];
      }
      s1 *= (1.0f / C);
      s2 *= (1.0f / C);
#pragma unroll
      for (int c = 0; c < C; ++c) g[c] = live ? inv * (da[c] - s1 - y[c] * s2) : 0.f;
    }
    {
      // depthwise-conv gradients: dw[t][c] += g[c] x[tok + t - 3][c], dwb[c] += g[c]
      const int ch = vec_reduce_channel<C>(lane);
      const bool lead = (lane % (32 / C)) == 0;
#pragma unroll
      for (int t = 0; t < 7; ++t) {
        const int ll = l + t - 3;
        const float mt = (ll >= 0 && ll < L) ? m : 0.f;
        const float* row = sx + (threadIdx.x + t) * RS;
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = mt * g[c] * row[c];
        const float r = warp_vec_reduce<C>(v, lane);
        if constexpr (REG_CH) rch[t] += r;
        else if (lead) atomicAdd(&sacc[t * C + ch], r);
      }
      float v[C];
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = m * g[c];
      const float r = warp_vec_reduce<C>(v, lane);
      if constexpr (REG_CH) rch[7] += r;
      else if (lead) atomicAdd(&sacc[7 * C + ch], r);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) sg[threadIdx.x * RS + c] = g[c];
    __syncthreads();
    // dX for the inner tokens
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
          const float* gr = sg + (threadIdx.x - t + 3) * RS;
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] = fmaf(sp[Lay::DW + t * C + c], gr[c], acc[c]);
        }
      }
      float4* dst = reinterpret_cast<float4*>(dX + static_cast<size_t>(tok) * C);
#pragma unroll
      for (int q = 0; q < V; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
    // weight gradients of the two pointwise convs: token-reduced outer products, 4 x 4 outputs per thread
    if constexpr (!REG_WG) {
      if (threadIdx.x < NTILE) {
        outer_tile<TH, TC, HS, CS>(sdu, sa, my_r0, my_c0, accW1);    // dW1[h][c]  += du[h] a[c]
        outer_tile<TH, TC, HS, CS>(sgl, sdg, my_r0, my_c0, accW2);   // dW2t[h][c] += gelu(u)[h] dg2[c]
      }
    }
  }
  if constexpr (REG_WG) {
    static_assert(!REG_WG || RW == 32, "one butterfly per matrix");
    const float r1 = warp_vec_reduce<RW>(regW1, lane), r2 = warp_vec_reduce<RW>(regW2, lane);
    atomicAdd(gparams + Lay::W1 + lane, r1);     // flat index h * C + c == lane
    atomicAdd(gparams + Lay::W2T + lane, r2);
  } else if (threadIdx.x < NTILE) {
#pragma unroll
    for (int i = 0; i < TH; ++i)
#pragma unroll
      for (int j = 0; j < TC; ++j) {
        atomicAdd(gparams + Lay::W1 + (my_r0 + i) * C + my_c0 + j, accW1[i][j]);
        atomicAdd(gparams + Lay::W2T + (my_r0 + i) * C + my_c0 + j, accW2[i][j]);
      }
  }
  if constexpr (REG_CH) {
    if ((lane % (32 / C)) == 0) {
      const int ch = vec_reduce_channel<C>(lane);
#pragma unroll
      for (int k = 0; k < 12; ++k) atomicAdd(&sacc[k * C + ch], rch[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 12 * C + H; i += SBB_THREADS) {
    // sacc order: dw[7][C] | dwb | lnw | lnb | b2 | gamma | b1[H]  ->  SmallBlockLayout offsets
    const int seg = i / C, c = i % C;
    const int off = seg < 7 ? Lay::DW + i : seg == 7 ? Lay::DWB + c : seg == 8 ? Lay::LNW + c : seg == 9 ? Lay::LNB + c
                  : seg == 10 ? Lay::B2 + c : seg == 11 ? Lay::GAMMA + c : Lay::B1 + (i - 12 * C);
    atomicAdd(gparams + off, sacc[i]);
  }
}

// ------------------------------------------------------------------------------------------ small Downsample backward
// Forward (downsample_small_kernel): n = LN(x) of two adjacent tokens (2*CIN values), y = W n + b.
// dY [M_out, COUT] -> dX [2*M_out, CIN]; parameter gradients in the SmallDownLayout image.
constexpr int SDB_THREADS = 128;
template <int CIN>
struct SmallDownBwd {
  static constexpr int COUT = 2 * CIN, K = 2 * CIN;
  static constexpr int T = (CIN >= 16) ? 4 : (CIN == 8 ? 2 : 1);          // register tile T x T
  static constexpr int TILES = (COUT / T) * (K / T);
  static constexpr int PER = (TILES + SDB_THREADS - 1) / SDB_THREADS;     // tiles per thread (2 for CIN = 32)
  static constexpr int OS = COUT + 4, KS = K + 4;                         // bf16 row strides
  static constexpr size_t BYTES = static_cast<size_t>(SmallDownLayout<CIN>::TOTAL) * 4 + static_cast<size_t>(SDB_THREADS) * (OS + KS) * 2;
};
template <int CIN>
constexpr size_t small_down_bwd_smem() { return SmallDownBwd<CIN>::BYTES; }

template <int CIN>
__global__ void __launch_bounds__(SDB_THREADS) downsample_small_bwd_kernel(const float* X, const float* dY, float* dX, int M_out,
                                                                           const float* __restrict__ params, float* __restrict__ gparams) {
  using Lay = SmallDownLayout<CIN>;
  using SD = SmallDownBwd<CIN>;
  constexpr int COUT = Lay::COUT, K = 2 * CIN, T = SD::T, OS = SD::OS, KS = SD::KS;
  extern __shared__ __align__(16) float smem_f[];
  float* sp = smem_f;
  __nv_bfloat16* sdy = reinterpret_cast<__nv_bfloat16*>(sp + Lay::TOTAL);   // [tok][OS]
  __nv_bfloat16* sn = sdy + SDB_THREADS * OS;                               // [tok][KS]
  float accW[SD::PER][T][T];
#pragma unroll
  for (int p = 0; p < SD::PER; ++p)
#pragma unroll
    for (int i = 0; i < T; ++i)
#pragma unroll
      for (int j = 0; j < T; ++j) accW[p][i][j] = 0.f;
  float glw[CIN], glb[CIN], gbo = 0.f;
#pragma unroll
  for (int c = 0; c < CIN; ++c) { glw[c] = 0.f; glb[c] = 0.f; }
  for (int i = threadIdx.x; i < Lay::TOTAL; i += SDB_THREADS) sp[i] = __ldg(params + i);
  const int ntiles = (M_out + SDB_THREADS - 1) / SDB_THREADS;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    const int tok = tile * SDB_THREADS + threadIdx.x;
    if (tok < M_out) {
      float n[K], dn[K], inv[2];
      const float4* src = reinterpret_cast<const float4*>(X + static_cast<size_t>(tok) * K);
#pragma unroll
      for (int q = 0; q < K / 4; ++q) {
        const float4 v = src[q];
        n[4 * q] = v.x; n[4 * q + 1] = v.y; n[4 * q + 2] = v.z; n[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float mean = 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) mean += n[t * CIN + c];
        mean *= (1.0f / CIN);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) { n[t * CIN + c] -= mean; var += n[t * CIN + c] * n[t * CIN + c]; }
        inv[t] = rsqrtf(var * (1.0f / CIN) + kLnEps);
#pragma unroll
        for (int c = 0; c < CIN; ++c) n[t * CIN + c] *= inv[t];   // xhat
      }
#pragma unroll
      for (int k = 0; k < K; k += 2) {
        dn[k] = 0.f;
        dn[k + 1] = 0.f;
        *reinterpret_cast<uint32_t*>(sn + threadIdx.x * KS + k) =
            pack_bf16x2(n[k] * sp[Lay::LNW + (k % CIN)] + sp[Lay::LNB + (k % CIN)],
                        n[k + 1] * sp[Lay::LNW + ((k + 1) % CIN)] + sp[Lay::LNB + ((k + 1) % CIN)]);
      }
      const float4* dsrc = reinterpret_cast<const float4*>(dY + static_cast<size_t>(tok) * COUT);
#pragma unroll 1
      for (int q = 0; q < COUT / 4; ++q) {
        const float4 v = dsrc[q];
        const float dy[4] = {v.x, v.y, v.z, v.w};
        *reinterpret_cast<uint2*>(sdy + threadIdx.x * OS + 4 * q) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int k = 0; k < K; ++k) dn[k] = fmaf(sp[Lay::W + (4 * q + u) * K + k], dy[u], dn[k]);
        }
      }
      float* dst = dX + static_cast<size_t>(tok) * K;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float d = dn[t * CIN + c];
          glw[c] += d * n[t * CIN + c];
          glb[c] += d;
          dn[t * CIN + c] = d * sp[Lay::LNW + c];
          s1 += dn[t * CIN + c];
          s2 += dn[t * CIN + c] * n[t * CIN + c];
        }
        s1 *= (1.0f / CIN);
        s2 *= (1.0f / CIN);
#pragma unroll
        for (int c = 0; c < CIN; ++c) dn[t * CIN + c] = inv[t] * (dn[t * CIN + c] - s1 - n[t * CIN + c] * s2);
      }
#pragma unroll
      for (int q = 0; q < K / 4; ++q)
        reinterpret_cast<float4*>(dst)[q] = make_float4(dn[4 * q], dn[4 * q + 1], dn[4 * q + 2], dn[4 * q + 3]);
    } else {
      for (int k = 0; k < K; k += 2) *reinterpret_cast<uint32_t*>(sn + threadIdx.x * KS + k) = 0u;
      for (int o = 0; o < COUT; o += 2) *reinterpret_cast<uint32_t*>(sdy + threadIdx.x * OS + o) = 0u;
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < SD::PER; ++p) {
      const int it = threadIdx.x + p * SDB_THREADS;
      if (it < SD::TILES) outer_tile<T, T, OS, KS>(sdy, sn, (it / (K / T)) * T, (it % (K / T)) * T, accW[p]);
    }
    if (threadIdx.x < COUT) {
      float s = 0.f;
      for (int t = 0; t < SDB_THREADS; ++t) s += __bfloat162float(sdy[t * OS + threadIdx.x]);
      gbo += s;
    }
  }
#pragma unroll
  for (int p = 0; p < SD::PER; ++p) {
    const int it = threadIdx.x + p * SDB_THREADS;
    if (it < SD::TILES) {
      const int r0 = (it / (K / T)) * T, c0 = (it % (K / T)) * T;
#pragma unroll
      for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; ++j) atomicAdd(gparams + Lay::W + (r0 + i) * K + c0 + j, accW[p][i][j]);
    }
  }
  if (threadIdx.x < COUT) atomicAdd(gparams + Lay::B + threadIdx.x, gbo);
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < CIN; ++c) {
    const float v1 = warp_sum(glw[c]), v2 = warp_sum(glb[c]);
    if (lane == 0) {
      atomicAdd(gparams + Lay::LNW + c, v1);
      atomicAdd(gparams + Lay::LNB + c, v2);
    }
  }
}

// ------------------------------------------------------------------------------------------ Stem (training variant + backward)
// (forward: stem_kernel in cnn_kernels.cuh, same parameter image)
__global__ void __launch_bounds__(256) stem_bwd_kernel(const float* __restrict__ audio, const float* __restrict__ dOut, int n_samples,
                                                       int L0, int total_tokens, const float* __restrict__ params,
                                                       float* __restrict__ gparams) {
  __shared__ float sp[STEM_P];
  __shared__ float sred[8][STEM_P];
  if (threadIdx.x < STEM_P) sp[threadIdx.x] = params[threadIdx.x];
  __syncthreads();
  float acc[STEM_P];
#pragma unroll
  for (int i = 0; i < STEM_P; ++i) acc[i] = 0.f;
  for (int tok = blockIdx.x * 256 + threadIdx.x; tok < total_tokens; tok += gridDim.x * 256) {
    const int b = tok / L0, l = tok - b * L0;
    const float* a0 = audio + static_cast<size_t>(b) * 2 * n_samples + static_cast<size_t>(l) * 5;
    const float* a1 = a0 + n_samples;
    float x[10];
#pragma unroll
    for (int k = 0; k < 5; ++k) { x[k] = __ldg(a0 + k); x[5 + k] = __ldg(a1 + k); }
    float y[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float a = sp[40 + o];
#pragma unroll
      for (int k = 0; k < 10; ++k) a = fmaf(sp[o * 10 + k], x[k], a);
      y[o] = a;
    }
    const float mean = 0.25f * (y[0] + y[1] + y[2] + y[3]);
    float var = 0.f;
#pragma unroll
    for (int o = 0; o < 4; ++o) { y[o] -= mean; var += y[o] * y[o]; }
    const float inv = rsqrtf(0.25f * var + kLnEps);
    const float4 d4 = reinterpret_cast<const float4*>(dOut)[tok];
    float d[4] = {d4.x, d4.y, d4.z, d4.w};
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      y[o] *= inv;
      acc[44 + o] += d[o] * y[o];
      acc[48 + o] += d[o];
      d[o] *= sp[44 + o];
      s1 += d[o];
      s2 += d[o] * y[o];
    }
    s1 *= 0.25f;
    s2 *= 0.25f;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float g = inv * (d[o] - s1 - y[o] * s2);
      acc[40 + o] += g;
#pragma unroll
      for (int k = 0; k < 10; ++k) acc[o * 10 + k] = fmaf(g, x[k], acc[o * 10 + k]);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < STEM_P; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) sred[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < STEM_P) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sred[w][threadIdx.x];
    atomicAdd(gparams + threadIdx.x, t);
  }
}

// ------------------------------------------------------------------------------------------ packing / unpacking / optimiser
// Packed weight images from the fp32 master parameters (reference leaf layout): element i of the flat packed list
// is master[src[i]] (* master[mul[i]] when mul[i] >= 0), or 0 when src[i] < 0, stored as fp32 or bf16 at byte offset
// dst_off[i] & 0x7fffffff of the arena (top bit = bf16).
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ master, const int* __restrict__ src,
                                                           const int* __restrict__ mul, const uint32_t* __restrict__ dst_off,
                                                           uint8_t* __restrict__ arena, size_t n) {
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) {
    const int s = src[i];
    float v = 0.f;
    if (s >= 0) {
      v = master[s];
      const int m = mul[i];
      if (m >= 0) v *= master[m];
    }
    const uint32_t d = dst_off[i];
    if (d & 0x80000000u) *reinterpret_cast<__nv_bfloat16*>(arena + (d & 0x7fffffffu)) = op1_rn(v);
    else *reinterpret_cast<float*>(arena + d) = v;
  }
}
// grads[src[i]] += gpack[i]: every master element receives contributions only from the packed copies the backward
// kernels wrote (the others stay zero), so plain atomics over a handful of aliases are enough.
__global__ void __launch_bounds__(256) grad_unpack_kernel(const float* __restrict__ gpack, const int* __restrict__ src,
                                                          float* __restrict__ grads, size_t n) {
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) {
    const float g = gpack[i];
    const int s = src[i];
    if (s >= 0 && g != 0.f) atomicAdd(grads + s, g);
  }
}

// The same for the destinations lo <= src[i] < hi only, with the gradient blob reached through a device pointer slot
// (rewritten by every a2m_backward, so the captured backward graph does not bake the caller's pointer in).
__global__ void __launch_bounds__(256) grad_unpack_range_kernel(const float* __restrict__ gpack, const int* __restrict__ src,
                                                                float* const* __restrict__ grads_slot, size_t n, int lo, int hi) {
  float* grads = *grads_slot;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) {
    const int s = src[i];
    if (s >= lo && s < hi) {
      const float g = gpack[i];
      if (g != 0.f) atomicAdd(grads + s, g);
    }
  }
}
__global__ void set_ptr_kernel(float** slot, float* p) { *slot = p; }

// Chunked variants: only the packed tensors the backward kernels actually write (about a third of the packed elements --
// every weight also exists as transposed / folded / re-laid-out copies that never receive gradients) are zeroed before and
// scattered after the backward.  chunks[i] = (first packed element, count <= 1024).
__global__ void __launch_bounds__(256) grad_zero_chunks_kernel(float* __restrict__ gpack, const int2* __restrict__ chunks, int n_chunks) {
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int2 ch = chunks[c];
    for (int i = threadIdx.x; i < ch.y; i += 256) gpack[ch.x + i] = 0.f;
  }
}
__global__ void __launch_bounds__(256) grad_unpack_chunks_kernel(const float* __restrict__ gpack, const int* __restrict__ src,
                                                                 float* const* __restrict__ grads_slot, const int2* __restrict__ chunks,
                                                                 int n_chunks, int lo, int hi) {
  float* grads = *grads_slot;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int2 ch = chunks[c];
    for (int i = threadIdx.x; i < ch.y; i += 256) {
      const int s = src[ch.x + i];
      if (s >= lo && s < hi) {
        const float g = gpack[ch.x + i];
        if (g != 0.f) atomicAdd(grads + s, g);
      }
    }
  }
}

// AdamW (optax.adamw as configured in train.py:646-726) followed by optax.clip_by_global_norm on the UPDATES
// (train.py:726 chains the clip after the optimiser).  Pass 1: moments, raw update u, sum u^2 and a finite flag.
// Pass 2: p += u * min(1, clip / ||u||).
struct AdamArgs {
  float lr, b1, b2, eps, wd, bc1, bc2, inv_div;   // bc = 1 - beta^t ; grads are multiplied by inv_div first
};
// Pass 0: stats[1] = number of non-finite gradient entries (train.py:320-322, grads_valid).  It runs BEFORE anything is
// written: with a non-finite gradient both passes below are no-ops, so the master weights and the moments are never
// poisoned (the reference restores a snapshot and halves the loss scale instead, train.py:369-377).
__global__ void __launch_bounds__(256) grad_finite_kernel(const float* __restrict__ g, size_t n, float* __restrict__ stats) {
  float bad = 0.f;
  if (reinterpret_cast<uintptr_t>(g) & 15) {   // unaligned blob (a caller's slice): scalar loads
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull)
      if (!isfinite(g[i])) bad += 1.f;
    bad = warp_sum(bad);
    if ((threadIdx.x & 31) == 0 && bad != 0.f) atomicAdd(stats + 1, bad);
    return;
  }
  const size_t n4 = n / 4;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n4; i += gridDim.x * 256ull) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    // x * 0 is 0 for every finite x and NaN for inf / NaN
    const float t = (v.x * 0.f + v.y * 0.f) + (v.z * 0.f + v.w * 0.f);
    if (t != 0.f) bad += 1.f;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    if (!isfinite(g[n4 * 4 + threadIdx.x])) bad += 1.f;
  }
  bad = warp_sum(bad);
  if ((threadIdx.x & 31) == 0 && bad != 0.f) atomicAdd(stats + 1, bad);
}
__global__ void __launch_bounds__(256) adamw_pass1_kernel(const float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                          float* __restrict__ v, float* __restrict__ u, const float* __restrict__ lr_mult,
                                                          size_t n, AdamArgs a, float* __restrict__ stats /* [0] sum u^2, [1] non-finite count */) {
  if (stats[1] != 0.f) return;   // written by grad_finite_kernel, earlier in the stream
  float ss = 0.f;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) {
    const float gi = g[i] * a.inv_div;
    const float mi = a.b1 * m[i] + (1.f - a.b1) * gi;
    const float vi = a.b2 * v[i] + (1.f - a.b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float mh = mi / a.bc1, vh = vi / a.bc2;
    const float lr = a.lr * (lr_mult ? lr_mult[i] : 1.f);
    const float ui = -lr * (mh / (sqrtf(vh) + a.eps) + a.wd * p[i]);
    u[i] = ui;
    ss += ui * ui;
  }
  ss = warp_sum(ss);
  __shared__ float s0[8];
  if ((threadIdx.x & 31) == 0) s0[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f;
    for (int i = 0; i < 8; ++i) t0 += s0[i];
    atomicAdd(stats, t0);
  }
}
__global__ void __launch_bounds__(256) adamw_pass2_kernel(float* __restrict__ p, const float* __restrict__ u, size_t n, float clip,
                                                          const float* __restrict__ stats) {
  if (stats[1] != 0.f) return;
  const float norm = sqrtf(stats[0]);
  const float f = (norm > clip) ? clip / norm : 1.f;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) p[i] += u[i] * f;
}

}  // namespace a2m
