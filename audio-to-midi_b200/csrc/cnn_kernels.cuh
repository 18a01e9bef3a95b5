// CUDA-core kernels of the ConvNeXt-style frontend (model.py:84-167, 756-759).
//
// Internal activation layout is TOKEN-MAJOR: X[b * L + l][c] fp32 (the reference's (C, L) transposed),
// so LayerNorm over channels is a contiguous row reduction and every 1x1 conv is a plain GEMM.
// C * L = 64 000 elements per window at every stage.
//
//   stem_kernel              Stem: conv k5 s5 (2->4) + LN(4)                       model.py:98-100
//   block_small_kernel<C>    whole Block for C in {4, 8, 16, 32} on CUDA cores      model.py:160-167
//   downsample_small_kernel  LN(Cin) + conv k2 s2 for Cin in {4, 8, 16, 32}         model.py:116-118
//   dwconv_ln_kernel<C>      depthwise k7 + LN -> bf16 GEMM operand, C in {64,128,256}  model.py:161-162
//   ln_rows_kernel<C>        LN over a row -> bf16 and/or fp32 (Downsample norm for Cin >= 64,
//                            final norm model.py:759, transformer/decoder norms model.py:190,539,546)
// All are HBM/L2-bound glue: coalesced 128-bit accesses, warp-shuffle reductions, no atomics.
//
// Activation pointers are deliberately NOT `__restrict__` and never read through __ldg: with programmatic
// dependent launch the previous kernel may still be writing them when this kernel starts, and ptxas hoists
// read-only (ld.global.nc) loads above griddepcontrol.wait (seen in SASS: LDG.E.CONSTANT before ACQBULK).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace a2m {

constexpr float kLnEps = 1e-5f;

// jax.nn.gelu (tanh form) through the hardware tanh: one MUFU op and 5 FMA-pipe ops per element.  The narrow stages spend
// most of their instructions here (2C activations per token, every stage), so the ex2 + rcp form cost 2x the issue slots.
__device__ __forceinline__ float gelu_tanh_cc(float x) {
  const float x2 = x * x;
  const float u = x * fmaf(0.0356774081f, x2, 0.7978845608f);   // sqrt(2/pi) (x + 0.044715 x^3)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// Sums v[i] over the 32 lanes for N (4 or 8) values at once and returns every total to every lane (in v): a transposing
// butterfly (N/2 + ... + 1 exchanges that halve the vector, then plain butterflies) leaves total i in lanes (32/N) i ..,
// then N broadcasts: 17 (N = 8) shuffles instead of the 40 dependent ones of eight warp_sum() calls.
template <int N>
__device__ __forceinline__ void warp_sum_all(float (&v)[N], int lane) {
  static_assert(N == 4 || N == 8, "warp_sum_all: 4 or 8 values");
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? v[i] : v[i + n / 2];
      const float keep = upper ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  float r = v[0];
#pragma unroll
  for (int o = (32 / N) / 2; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = __shfl_sync(0xffffffffu, r, (32 / N) * i);
}
__device__ __forceinline__ void warp_sum8_all(float (&v)[8], int lane) { warp_sum_all<8>(v, lane); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------ Stem
// Packed fp32 parameter image in the weights arena: w[4][2][5] (40) | b[4] | lnw[4] | lnb[4].  Read from device memory so
// that a handle whose weights are re-packed after every optimizer step (training) never runs the stem on stale values.
constexpr int STEM_P = 52;
// audio [B, 2, n_samples] fp32 or IEEE binary16 -> X [B * L0, 4] fp32, L0 = n_samples / 5.  One thread per output token.
// The f16 input is lossless for audio that went through load_full_audio: it rounds every sample to f16 (python.rs:235-264),
// so a host caller can ship half the bytes (a2m_submit_host_ex).
__device__ __forceinline__ float stem_ld(const float* p) { return __ldg(p); }
__device__ __forceinline__ float stem_ld(const __half* p) { return __half2float(__ldg(p)); }
template <class T>
__global__ void __launch_bounds__(256) stem_kernel(const T* __restrict__ audio, float* __restrict__ out, int n_samples,
                                                   int L0, int total_tokens, const float* __restrict__ params) {
  __shared__ float sp[STEM_P];
  if (threadIdx.x < STEM_P) sp[threadIdx.x] = params[threadIdx.x];
  __syncthreads();
  const int tok = blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= total_tokens) return;
  const int b = tok / L0, l = tok - b * L0;
  const T* a0 = audio + static_cast<size_t>(b) * 2 * n_samples + static_cast<size_t>(l) * 5;
  const T* a1 = a0 + n_samples;
  float x[10];
#pragma unroll
  for (int k = 0; k < 5; ++k) { x[k] = stem_ld(a0 + k); x[5 + k] = stem_ld(a1 + k); }
  float y[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    float acc = sp[40 + o];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc = fmaf(sp[o * 10 + k], x[k], acc);
    y[o] = acc;
  }
  const float mean = 0.25f * (y[0] + y[1] + y[2] + y[3]);
  float var = 0.f;
#pragma unroll
  for (int o = 0; o < 4; ++o) var += (y[o] - mean) * (y[o] - mean);
  const float inv = rsqrtf(0.25f * var + kLnEps);
  float4 r;
  r.x = (y[0] - mean) * inv * sp[44] + sp[48];
  r.y = (y[1] - mean) * inv * sp[45] + sp[49];
  r.z = (y[2] - mean) * inv * sp[46] + sp[50];
  r.w = (y[3] - mean) * inv * sp[47] + sp[51];
  reinterpret_cast<float4*>(out)[tok] = r;
}

// ------------------------------------------------------------------------------------------ small Block
// Packed fp32 parameter image of one Block for the CUDA-core kernel (built on the host):
//   dw[7][C] | dwb[C] | lnw[C] | lnb[C] | w1[H][C] | b1[H] | w2t[H][C] (= point_conv_2 transposed) | b2[C] | gamma[C]
template <int C>
struct SmallBlockLayout {
  static constexpr int H = 2 * C;
  static constexpr int DW = 0;
  static constexpr int DWB = DW + 7 * C;
  static constexpr int LNW = DWB + C;
  static constexpr int LNB = LNW + C;
  static constexpr int W1 = LNB + C;
  static constexpr int B1 = W1 + H * C;
  static constexpr int W2T = B1 + H;
  static constexpr int B2 = W2T + H * C;
  static constexpr int GAMMA = B2 + C;
  static constexpr int TOTAL = GAMMA + C;
};

constexpr int SB_TOK = 128;  // threads per CTA
// Tokens per thread (a CTA covers SB_TOK * small_block_tpt<C>() consecutive tokens).
template <int C>
__host__ __device__ constexpr int small_block_tpt() { return 1; }   // measured: 8 / 4 tokens per thread at C = 4 / 8 were 30 % SLOWER (fewer resident warps)

template <int C>
__global__ void __launch_bounds__(SB_TOK) block_small_kernel(const float* Xin, float* Xout,
                                                             int L, int M, const float* __restrict__ params) {
  using Lay = SmallBlockLayout<C>;
  constexpr int H = Lay::H;
  constexpr int TPT = small_block_tpt<C>();
  constexpr int TILE = SB_TOK * TPT;
  constexpr int RS = (C == 4) ? 4 : C + 4;  // padded row stride (floats): conflict-free 128-bit row reads
  extern __shared__ __align__(16) float smem_f[];
  float* sp = smem_f;                          // parameters
  float* sx = smem_f + ((Lay::TOTAL + 3) & ~3);  // (TILE + 6) rows of input

  pdl_launch_dependents();
  static_assert(Lay::TOTAL % 4 == 0, "parameter image is copied as 16-byte vectors");
  copy_const_to_smem<Lay::TOTAL / 4, SB_TOK>(sp, params, threadIdx.x);
  pdl_wait();  // parameters are constants; activations of the previous kernel are read below
  const int tile0 = blockIdx.x * TILE;
  // rows tile0-3 .. tile0+TILE+2, zero outside [0, M): every load of a thread is issued before its first store
  constexpr int V = C / 4;
  {
    constexpr int NV = (TILE + 6) * V;
    constexpr int PER = (NV + SB_TOK - 1) / SB_TOK;
    float4 v[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = threadIdx.x + k * SB_TOK;
      const int r = i / V, q = i - r * V;
      const int g = tile0 - 3 + r;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < NV && g >= 0 && g < M) v[k] = reinterpret_cast<const float4*>(Xin + static_cast<size_t>(g) * C)[q];   // plain load: never hoisted above pdl_wait
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = threadIdx.x + k * SB_TOK;
      const int r = i / V, q = i - r * V;
      if (i < NV) reinterpret_cast<float4*>(sx + r * RS)[q] = v[k];
    }
  }
  __syncthreads();

#pragma unroll 1
  for (int it = 0; it < TPT; ++it) {
  const int trow = it * SB_TOK + threadIdx.x;   // row of this token inside the tile
  const int tok = tile0 + trow;
  if (tok >= M) return;
  const int l = tok % L;

  // depthwise conv k7, zero "SAME" padding at the WINDOW boundary (not the tile/batch boundary)
  float y[C];
#pragma unroll
  for (int c = 0; c < C; ++c) y[c] = sp[Lay::DWB + c];
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    const int ll = l + t - 3;
    if (ll >= 0 && ll < L) {
      const float* row = sx + (trow + t) * RS;
#pragma unroll
      for (int q = 0; q < V; ++q) {
        const float4 xv = reinterpret_cast<const float4*>(row)[q];
        const float4 wv = reinterpret_cast<const float4*>(sp + Lay::DW + t * C)[q];
        y[4 * q] = fmaf(wv.x, xv.x, y[4 * q]);
        y[4 * q + 1] = fmaf(wv.y, xv.y, y[4 * q + 1]);
        y[4 * q + 2] = fmaf(wv.z, xv.z, y[4 * q + 2]);
        y[4 * q + 3] = fmaf(wv.w, xv.w, y[4 * q + 3]);
      }
    }
  }
  // LayerNorm over channels (fp32, biased variance)
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) mean += y[c];
  mean *= (1.0f / C);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) var += (y[c] - mean) * (y[c] - mean);
  const float inv = rsqrtf(var * (1.0f / C) + kLnEps);
#pragma unroll
  for (int c = 0; c < C; ++c) y[c] = (y[c] - mean) * inv * sp[Lay::LNW + c] + sp[Lay::LNB + c];

  // pointwise C -> H, GELU, pointwise H -> C, one hidden unit at a time (weights broadcast from smem)
  float o[C];
#pragma unroll
  for (int c = 0; c < C; ++c) o[c] = sp[Lay::B2 + c];
#pragma unroll 2
  for (int h = 0; h < H; ++h) {
    float a = sp[Lay::B1 + h];
    const float4* w1 = reinterpret_cast<const float4*>(sp + Lay::W1 + h * C);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      const float4 w = w1[q];
      a = fmaf(w.x, y[4 * q], a);
      a = fmaf(w.y, y[4 * q + 1], a);
      a = fmaf(w.z, y[4 * q + 2], a);
      a = fmaf(w.w, y[4 * q + 3], a);
    }
    const float gl = gelu_tanh_cc(a);
    const float4* w2 = reinterpret_cast<const float4*>(sp + Lay::W2T + h * C);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      const float4 w = w2[q];
      o[4 * q] = fmaf(w.x, gl, o[4 * q]);
      o[4 * q + 1] = fmaf(w.y, gl, o[4 * q + 1]);
      o[4 * q + 2] = fmaf(w.z, gl, o[4 * q + 2]);
      o[4 * q + 3] = fmaf(w.w, gl, o[4 * q + 3]);
    }
  }
  // layer scale + residual.  Out of place: neighbouring CTAs read rows of this tile as their halo.
  const float* xr = sx + (trow + 3) * RS;
  float* dst = Xout + static_cast<size_t>(tok) * C;
#pragma unroll
  for (int q = 0; q < V; ++q) {
    float4 r;
    r.x = fmaf(sp[Lay::GAMMA + 4 * q], o[4 * q], xr[4 * q]);
    r.y = fmaf(sp[Lay::GAMMA + 4 * q + 1], o[4 * q + 1], xr[4 * q + 1]);
    r.z = fmaf(sp[Lay::GAMMA + 4 * q + 2], o[4 * q + 2], xr[4 * q + 2]);
    r.w = fmaf(sp[Lay::GAMMA + 4 * q + 3], o[4 * q + 3], xr[4 * q + 3]);
    reinterpret_cast<float4*>(dst)[q] = r;
  }
  }
}


// ------------------------------------------------------------------------------------------ three small Blocks fused
// The three Blocks of a narrow stage (C in {4, 8}: a 128-token tile is 2-4 KB, every launch is a latency-bound chain
// of load -> barrier -> ~300 instructions per token -> store) in ONE launch: the tile is loaded once with a halo of 9
// tokens, Block k is evaluated by one thread per token for the rows the next Block still needs (halo shrinking by 3 per
// Block, ping-pong between two shared-memory buffers), the third Block's SS_TOK rows are stored.  One third of the global
// traffic and of the launches of three block_small_kernel calls; the arithmetic per token is identical (same device
// function), so are the results.
template <int C>
__device__ __forceinline__ void small_block_token(const float* __restrict__ sp, const float* rows /* row of token-3 */, int RS,
                                                  int l, int L, float* out /* [C] */) {
  using Lay = SmallBlockLayout<C>;
  constexpr int H = Lay::H, V = C / 4;
  float y[C];
#pragma unroll
  for (int c = 0; c < C; ++c) y[c] = sp[Lay::DWB + c];
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    const int ll = l + t - 3;
    if (ll >= 0 && ll < L) {   // zero "SAME" padding at the WINDOW boundary
      const float* row = rows + t * RS;
#pragma unroll
      for (int q = 0; q < V; ++q) {
        const float4 xv = reinterpret_cast<const float4*>(row)[q];
        const float4 wv = reinterpret_cast<const float4*>(sp + Lay::DW + t * C)[q];
        y[4 * q] = fmaf(wv.x, xv.x, y[4 * q]);
        y[4 * q + 1] = fmaf(wv.y, xv.y, y[4 * q + 1]);
        y[4 * q + 2] = fmaf(wv.z, xv.z, y[4 * q + 2]);
        y[4 * q + 3] = fmaf(wv.w, xv.w, y[4 * q + 3]);
      }
    }
  }
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) mean += y[c];
  mean *= (1.0f / C);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) var += (y[c] - mean) * (y[c] - mean);
  const float inv = rsqrtf(var * (1.0f / C) + kLnEps);
#pragma unroll
  for (int c = 0; c < C; ++c) y[c] = (y[c] - mean) * inv * sp[Lay::LNW + c] + sp[Lay::LNB + c];
  float o[C];
#pragma unroll
  for (int c = 0; c < C; ++c) o[c] = sp[Lay::B2 + c];
#pragma unroll 2
  for (int h = 0; h < H; ++h) {
    float a = sp[Lay::B1 + h];
    const float4* w1 = reinterpret_cast<const float4*>(sp + Lay::W1 + h * C);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      const float4 w = w1[q];
      a = fmaf(w.x, y[4 * q], a);
      a = fmaf(w.y, y[4 * q + 1], a);
      a = fmaf(w.z, y[4 * q + 2], a);
      a = fmaf(w.w, y[4 * q + 3], a);
    }
    const float gl = gelu_tanh_cc(a);
    const float4* w2 = reinterpret_cast<const float4*>(sp + Lay::W2T + h * C);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      const float4 w = w2[q];
      o[4 * q] = fmaf(w.x, gl, o[4 * q]);
      o[4 * q + 1] = fmaf(w.y, gl, o[4 * q + 1]);
      o[4 * q + 2] = fmaf(w.z, gl, o[4 * q + 2]);
      o[4 * q + 3] = fmaf(w.w, gl, o[4 * q + 3]);
    }
  }
  const float* xr = rows + 3 * RS;   // layer scale + residual
#pragma unroll
  for (int c = 0; c < C; ++c) out[c] = fmaf(sp[Lay::GAMMA + c], o[c], xr[c]);
}

// Two ADJACENT tokens (buffer rows r and r + 1) per thread: every weight vector fetched from shared memory feeds both tokens and
// the 8 input rows of the two depthwise windows are read once instead of 2 x 7 -- these kernels spend as many issue slots on
// LDS as on arithmetic.  Each token's operations are those of small_block_token, in the same order: results are bit-identical.
// rows = row of (token 0) - 3; `two` = false: only token 0 exists (its partner's results are garbage the caller ignores).
template <int C>
__device__ __forceinline__ void small_block_pair(const float* __restrict__ sp, const float* rows, int RS, int l0, int l1, int L, bool two,
                                                 float* out0 /* [C] */, float* out1 /* [C] */) {
  using Lay = SmallBlockLayout<C>;
  constexpr int H = Lay::H, V = C / 4;
  float y0[C], y1[C];
#pragma unroll
  for (int c = 0; c < C; ++c) y0[c] = y1[c] = sp[Lay::DWB + c];
  float xa[C], xb[C];                      // rows t and t + 1 of the 8-row window
#pragma unroll
  for (int q = 0; q < V; ++q) {
    const float4 v = reinterpret_cast<const float4*>(rows)[q];
    xa[4 * q] = v.x; xa[4 * q + 1] = v.y; xa[4 * q + 2] = v.z; xa[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    if (t < 6 || two) {
#pragma unroll
      for (int q = 0; q < V; ++q) {
        const float4 v = reinterpret_cast<const float4*>(rows + (t + 1) * RS)[q];
        xb[4 * q] = v.x; xb[4 * q + 1] = v.y; xb[4 * q + 2] = v.z; xb[4 * q + 3] = v.w;
      }
    }
    float w[C];
#pragma unroll
    for (int q = 0; q < V; ++q) {
      const float4 v = reinterpret_cast<const float4*>(sp + Lay::DW + t * C)[q];
      w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
    const int ll0 = l0 + t - 3, ll1 = l1 + t - 3;   // zero "SAME" padding at the WINDOW boundary
    if (ll0 >= 0 && ll0 < L) {
#pragma unroll
      for (int c = 0; c < C; ++c) y0[c] = fmaf(w[c], xa[c], y0[c]);
    }
    if (ll1 >= 0 && ll1 < L) {
#pragma unroll
      for (int c = 0; c < C; ++c) y1[c] = fmaf(w[c], xb[c], y1[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) xa[c] = xb[c];
  }
  auto layer_norm = [&](float* y) {
    float mean = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) mean += y[c];
    mean *= (1.0f / C);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) var += (y[c] - mean) * (y[c] - mean);
    const float inv = rsqrtf(var * (1.0f / C) + kLnEps);
#pragma unroll
    for (int c = 0; c < C; ++c) y[c] = (y[c] - mean) * inv * sp[Lay::LNW + c] + sp[Lay::LNB + c];
  };
  layer_norm(y0);
  layer_norm(y1);
  float o0[C], o1[C];
#pragma unroll
  for (int c = 0; c < C; ++c) o0[c] = o1[c] = sp[Lay::B2 + c];
#pragma unroll 2
  for (int h = 0; h < H; ++h) {
    float a0 = sp[Lay::B1 + h], a1 = a0;
    const float4* w1 = reinterpret_cast<const float4*>(sp + Lay::W1 + h * C);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      const float4 w = w1[q];
      a0 = fmaf(w.x, y0[4 * q], a0);
      a0 = fmaf(w.y, y0[4 * q + 1], a0);
      a0 = fmaf(w.z, y0[4 * q + 2], a0);
      a0 = fmaf(w.w, y0[4 * q + 3], a0);
      a1 = fmaf(w.x, y1[4 * q], a1);
      a1 = fmaf(w.y, y1[4 * q + 1], a1);
      a1 = fmaf(w.z, y1[4 * q + 2], a1);
      a1 = fmaf(w.w, y1[4 * q + 3], a1);
    }
    const float g0 = gelu_tanh_cc(a0), g1 = gelu_tanh_cc(a1);
    const float4* w2 = reinterpret_cast<const float4*>(sp + Lay::W2T + h * C);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      const float4 w = w2[q];
      o0[4 * q] = fmaf(w.x, g0, o0[4 * q]);
      o0[4 * q + 1] = fmaf(w.y, g0, o0[4 * q + 1]);
      o0[4 * q + 2] = fmaf(w.z, g0, o0[4 * q + 2]);
      o0[4 * q + 3] = fmaf(w.w, g0, o0[4 * q + 3]);
      o1[4 * q] = fmaf(w.x, g1, o1[4 * q]);
      o1[4 * q + 1] = fmaf(w.y, g1, o1[4 * q + 1]);
      o1[4 * q + 2] = fmaf(w.z, g1, o1[4 * q + 2]);
      o1[4 * q + 3] = fmaf(w.w, g1, o1[4 * q + 3]);
    }
  }
  const float* xr0 = rows + 3 * RS;   // layer scale + residual
  const float* xr1 = rows + 4 * RS;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float g = sp[Lay::GAMMA + c];
    out0[c] = fmaf(g, o0[c], xr0[c]);
    out1[c] = fmaf(g, o1[c], xr1[c]);
  }
}

#ifndef A2M_SS_PAIR
#define A2M_SS_PAIR 1                    // two adjacent tokens per thread (small_block_pair): 0.107 -> 0.091 ms per step for the two stages
#endif
#ifndef A2M_SS_TOK
#define A2M_SS_TOK (A2M_SS_PAIR ? 256 : 128)
#endif
#ifndef A2M_SS_THREADS
#define A2M_SS_THREADS 160
#endif
constexpr int SS_TOK = A2M_SS_TOK;              // output tokens per CTA
constexpr int SS_THREADS = A2M_SS_THREADS;      // >= the rows (or row pairs) evaluated by the first Block: SS_TOK + 12
static_assert(SS_THREADS * (A2M_SS_PAIR ? 2 : 1) >= SS_TOK + 12, "one thread per row (pair) of the first Block");
template <int C>
struct SmallStageCfg {
  static constexpr int RS = (C == 4) ? 4 : C + 4;
  static constexpr int P = (SmallBlockLayout<C>::TOTAL + 3) & ~3;
  static constexpr int ROWS = SS_TOK + 18;
  static constexpr size_t SMEM = (3 * P + 2 * ROWS * RS) * sizeof(float);
};

// params0/1/2: SmallBlockLayout images of the stage's three Blocks.
template <int C>
__global__ void __launch_bounds__(SS_THREADS) stage_small_kernel(const float* Xin, float* Xout, int L, int M, const float* __restrict__ params0,
                                                                 const float* __restrict__ params1, const float* __restrict__ params2) {
  using Lay = SmallBlockLayout<C>;
  using Cfg = SmallStageCfg<C>;
  constexpr int RS = Cfg::RS, V = C / 4;
  static_assert(Lay::TOTAL % 4 == 0, "parameter images are copied as 16-byte vectors");
  extern __shared__ __align__(16) float smem_f[];
  float* sp = smem_f;                                   // three parameter images
  float* bufA = smem_f + 3 * Cfg::P;                    // rows tile0-9 .. tile0+SS_TOK+8
  float* bufB = bufA + Cfg::ROWS * RS;

  pdl_launch_dependents();
  copy_const_to_smem<Lay::TOTAL / 4, SS_THREADS>(sp, params0, threadIdx.x);
  copy_const_to_smem<Lay::TOTAL / 4, SS_THREADS>(sp + Cfg::P, params1, threadIdx.x);
  copy_const_to_smem<Lay::TOTAL / 4, SS_THREADS>(sp + 2 * Cfg::P, params2, threadIdx.x);
  pdl_wait();  // parameters are constants; activations of the previous kernel are read below (plain loads)
  const int tile0 = blockIdx.x * SS_TOK;
  {
    constexpr int NV = Cfg::ROWS * V, PER = (NV + SS_THREADS - 1) / SS_THREADS;
    float4 v[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = threadIdx.x + k * SS_THREADS;
      const int r = i / V, q = i - r * V;
      const int g = tile0 - 9 + r;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < NV && g >= 0 && g < M) v[k] = reinterpret_cast<const float4*>(Xin + static_cast<size_t>(g) * C)[q];
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = threadIdx.x + k * SS_THREADS;
      const int r = i / V, q = i - r * V;
      if (i < NV) reinterpret_cast<float4*>(bufA + r * RS)[q] = v[k];
    }
  }
  __syncthreads();
  // Block k (k = 0, 1, 2) produces buffer rows 3(k+1) .. ROWS - 3(k+1) - 1  (buffer row r <-> token tile0 - 9 + r)
  float* src = bufA;
  float* dst = bufB;
#pragma unroll 1
  for (int k = 0; k < 3; ++k) {
    const int first = 3 * (k + 1), count = Cfg::ROWS - 6 * (k + 1);
#if A2M_SS_PAIR
    if (2 * static_cast<int>(threadIdx.x) < count) {
      const int r = first + 2 * threadIdx.x;
      const int tok = tile0 - 9 + r;
      const bool two = 2 * static_cast<int>(threadIdx.x) + 1 < count;
      const bool v0 = tok >= 0 && tok < M, v1 = two && tok + 1 >= 0 && tok + 1 < M;
      float o0[C], o1[C];
      if (v0 || v1) {
        const int l0 = v0 ? tok % L : 0, l1 = v1 ? (tok + 1) % L : 0;
        small_block_pair<C>(sp + k * Cfg::P, src + (r - 3) * RS, RS, l0, l1, L, two, o0, o1);
      }
      if (!v0) {
#pragma unroll
        for (int c = 0; c < C; ++c) o0[c] = 0.f;
      }
      if (!v1) {
#pragma unroll
        for (int c = 0; c < C; ++c) o1[c] = 0.f;
      }
      if (k < 2) {
#pragma unroll
        for (int q = 0; q < V; ++q) reinterpret_cast<float4*>(dst + r * RS)[q] = make_float4(o0[4 * q], o0[4 * q + 1], o0[4 * q + 2], o0[4 * q + 3]);
        if (two) {
#pragma unroll
          for (int q = 0; q < V; ++q) reinterpret_cast<float4*>(dst + (r + 1) * RS)[q] = make_float4(o1[4 * q], o1[4 * q + 1], o1[4 * q + 2], o1[4 * q + 3]);
        }
      } else {
        if (v0) {
          float* g = Xout + static_cast<size_t>(tok) * C;
#pragma unroll
          for (int q = 0; q < V; ++q) reinterpret_cast<float4*>(g)[q] = make_float4(o0[4 * q], o0[4 * q + 1], o0[4 * q + 2], o0[4 * q + 3]);
        }
        if (v1) {
          float* g = Xout + static_cast<size_t>(tok + 1) * C;
#pragma unroll
          for (int q = 0; q < V; ++q) reinterpret_cast<float4*>(g)[q] = make_float4(o1[4 * q], o1[4 * q + 1], o1[4 * q + 2], o1[4 * q + 3]);
        }
      }
    }
#else
    if (static_cast<int>(threadIdx.x) < count) {
      const int r = first + threadIdx.x;
      const int tok = tile0 - 9 + r;
      float o[C];
      if (tok >= 0 && tok < M) {
        small_block_token<C>(sp + k * Cfg::P, src + (r - 3) * RS, RS, tok % L, L, o);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = 0.f;
      }
      if (k < 2) {
#pragma unroll
        for (int q = 0; q < V; ++q) reinterpret_cast<float4*>(dst + r * RS)[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      } else if (tok < M) {
        float* g = Xout + static_cast<size_t>(tok) * C;
#pragma unroll
        for (int q = 0; q < V; ++q) reinterpret_cast<float4*>(g)[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      }
    }
#endif
    __syncthreads();
    float* t = src; src = dst; dst = t;
  }
}

// ------------------------------------------------------------------------------------------ small Downsample
// Packed parameters: lnw[Cin] | lnb[Cin] | w[Cout][2*Cin] (k index = tap * Cin + c) | b[Cout]
template <int CIN>
struct SmallDownLayout {
  static constexpr int COUT = 2 * CIN;
  static constexpr int LNW = 0;
  static constexpr int LNB = CIN;
  static constexpr int W = 2 * CIN;
  static constexpr int B = W + COUT * 2 * CIN;
  static constexpr int TOTAL = B + COUT;
};

// X [M_in, CIN] fp32 -> Y [M_in / 2, 2*CIN] fp32.  One thread per OUTPUT token (two adjacent input tokens;
// L_in is even at every stage so a pair never straddles a window).
template <int CIN>
__global__ void __launch_bounds__(128) downsample_small_kernel(const float* X, float* Y,
                                                               int M_out, const float* __restrict__ params) {
  using Lay = SmallDownLayout<CIN>;
  constexpr int COUT = Lay::COUT;
  constexpr int K = 2 * CIN;
  extern __shared__ __align__(16) float smem_f[];
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < Lay::TOTAL; i += blockDim.x) smem_f[i] = __ldg(params + i);
  pdl_wait();
  __syncthreads();
  const int tok = blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= M_out) return;

  float n[K];
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
    for (int c = 0; c < CIN; ++c) var += (n[t * CIN + c] - mean) * (n[t * CIN + c] - mean);
    const float inv = rsqrtf(var * (1.0f / CIN) + kLnEps);
#pragma unroll
    for (int c = 0; c < CIN; ++c)
      n[t * CIN + c] = (n[t * CIN + c] - mean) * inv * smem_f[Lay::LNW + c] + smem_f[Lay::LNB + c];
  }
  float* dst = Y + static_cast<size_t>(tok) * COUT;
#pragma unroll 1
  for (int o = 0; o < COUT; o += 4) {
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float a = smem_f[Lay::B + o + u];
      const float4* w = reinterpret_cast<const float4*>(smem_f + Lay::W + (o + u) * K);
#pragma unroll
      for (int q = 0; q < K / 4; ++q) {
        const float4 wv = w[q];
        a = fmaf(wv.x, n[4 * q], a);
        a = fmaf(wv.y, n[4 * q + 1], a);
        a = fmaf(wv.z, n[4 * q + 2], a);
        a = fmaf(wv.w, n[4 * q + 3], a);
      }
      acc[u] = a;
    }
    reinterpret_cast<float4*>(dst)[o / 4] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// ------------------------------------------------------------------------------------------ lane <-> channel map
// A warp owns one row of C channels.  Lane `lane` holds G groups of VW contiguous channels:
//   channel(g, j) = g * 32 * VW + lane * VW + j,   VW = min(4, C / 32),  G = C / (32 * VW)
// so every warp-wide access is one contiguous 32*VW*4-byte segment (coalesced in global memory,
// bank-conflict free in shared memory).
template <int C>
struct RowMap {
  static constexpr int VW = (C / 32 >= 4) ? 4 : C / 32;
  static constexpr int G = C / (32 * VW);
  static constexpr int PER = VW * G;
  static_assert(C % 64 == 0 && VW >= 2, "row kernels need C in {64, 128, 256}");
  __device__ static __forceinline__ int chan(int lane, int g) { return g * 32 * VW + lane * VW; }
  __device__ static __forceinline__ void load(const float* row, int lane, float* x) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if constexpr (VW == 4) {
        const float4 v = *reinterpret_cast<const float4*>(row + chan(lane, g));
        x[4 * g] = v.x; x[4 * g + 1] = v.y; x[4 * g + 2] = v.z; x[4 * g + 3] = v.w;
      } else {
        const float2 v = *reinterpret_cast<const float2*>(row + chan(lane, g));
        x[2 * g] = v.x; x[2 * g + 1] = v.y;
      }
    }
  }
  __device__ static __forceinline__ void store_f32(float* row, int lane, const float* x) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if constexpr (VW == 4)
        *reinterpret_cast<float4*>(row + chan(lane, g)) = make_float4(x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3]);
      else
        *reinterpret_cast<float2*>(row + chan(lane, g)) = make_float2(x[2 * g], x[2 * g + 1]);
    }
  }
  __device__ static __forceinline__ void store_bf16(__nv_bfloat16* row, int lane, const float* x) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if constexpr (VW == 4) {
        __nv_bfloat162 a = op2_rn(x[4 * g], x[4 * g + 1]);
        __nv_bfloat162 b = op2_rn(x[4 * g + 2], x[4 * g + 3]);
        uint2 q;
        q.x = *reinterpret_cast<uint32_t*>(&a);
        q.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(row + chan(lane, g)) = q;
      } else {
        *reinterpret_cast<__nv_bfloat162*>(row + chan(lane, g)) = op2_rn(x[2 * g], x[2 * g + 1]);
      }
    }
  }
  // LayerNorm of the row held across the warp (fp32 statistics, biased variance, eps inside the sqrt)
  __device__ static __forceinline__ void layer_norm(float* x, const float* lw, const float* lb) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) s += x[j];
    const float mean = warp_sum(s) * (1.0f / C);
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) v += (x[j] - mean) * (x[j] - mean);
    const float inv = rsqrtf(warp_sum(v) * (1.0f / C) + kLnEps);
#pragma unroll
    for (int j = 0; j < PER; ++j) x[j] = (x[j] - mean) * inv * lw[j] + lb[j];
  }
};

// ------------------------------------------------------------------------------------------ dwconv + LN (C >= 64)
// Packed parameters: dw[7][C] | dwb[C] | lnw[C] | lnb[C]
constexpr int DW_TOK = 32;       // tokens per CTA
constexpr int DW_THREADS = 256;  // 8 warps, 4 tokens each

template <int C>
__global__ void __launch_bounds__(DW_THREADS) dwconv_ln_kernel(const float* X, __nv_bfloat16* A,
                                                               int L, int M, const float* __restrict__ params) {
  using RM = RowMap<C>;
  constexpr int PER = RM::PER;
  extern __shared__ __align__(16) float smem_f[];
  float* sx = smem_f;  // (DW_TOK + 6) x C
  pdl_launch_dependents();
  pdl_wait();
  const int tile0 = blockIdx.x * DW_TOK;
  constexpr int V = C / 4;
  for (int i = threadIdx.x; i < (DW_TOK + 6) * V; i += DW_THREADS) {
    const int r = i / V, q = i - r * V;
    const int g = tile0 - 3 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g >= 0 && g < M) v = reinterpret_cast<const float4*>(X + static_cast<size_t>(g) * C)[q];
    reinterpret_cast<float4*>(sx + r * C)[q] = v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float w[7][PER], bias[PER], lw[PER], lb[PER];
#pragma unroll
  for (int t = 0; t < 7; ++t) RM::load(params + t * C, lane, w[t]);
  RM::load(params + 7 * C, lane, bias);
  RM::load(params + 8 * C, lane, lw);
  RM::load(params + 9 * C, lane, lb);
  __syncthreads();

#pragma unroll 1
  for (int i = 0; i < DW_TOK / 8; ++i) {
    const int lt = warp * (DW_TOK / 8) + i;
    const int tok = tile0 + lt;
    if (tok >= M) break;
    const int l = tok % L;
    float y[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) y[j] = bias[j];
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      const int ll = l + t - 3;
      if (ll >= 0 && ll < L) {  // zero "SAME" padding at the window boundary
        float xr[PER];
        RM::load(sx + (lt + t) * C, lane, xr);
#pragma unroll
        for (int j = 0; j < PER; ++j) y[j] = fmaf(w[t][j], xr[j], y[j]);
      }
    }
    RM::layer_norm(y, lw, lb);
    RM::store_bf16(A + static_cast<size_t>(tok) * C, lane, y);
  }
}

// ------------------------------------------------------------------------------------------ row LayerNorm
// One warp per OUTPUT row of C channels: fp32 in -> bf16 (GEMM operand) and/or fp32 out.  Rows may be
// re-mapped from a compact [B * Lin] layout into a padded [B * Lout] layout (final CNN norm -> transformer
// buffers, T = 250 -> 256); the pad rows are written as zeros on every call so that everything later
// derived from them is finite and deterministic (0 * NaN in a masked P.V product would poison real rows).
template <int C>
__global__ void __launch_bounds__(256) ln_rows_kernel(const float* X, int out_rows, int Lin, int Lout,
                                                      const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                      __nv_bfloat16* __restrict__ out16, float* __restrict__ out32) {
  using RM = RowMap<C>;
  constexpr int PER = RM::PER;
  pdl_launch_dependents();
  pdl_wait();
  const int orow = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (orow >= out_rows) return;
  const int lane = threadIdx.x & 31;
  const int b = orow / Lout, t = orow - b * Lout;
  float x[PER];
  if (t < Lin) {
    float lw[PER], lb[PER];
    RM::load(X + (static_cast<size_t>(b) * Lin + t) * C, lane, x);
    RM::load(lnw, lane, lw);
    RM::load(lnb, lane, lb);
    RM::layer_norm(x, lw, lb);
  } else {
#pragma unroll
    for (int j = 0; j < PER; ++j) x[j] = 0.f;
  }
  if (out32 != nullptr) RM::store_f32(out32 + static_cast<size_t>(orow) * C, lane, x);
  if (out16 != nullptr) RM::store_bf16(out16 + static_cast<size_t>(orow) * C, lane, x);
}

}  // namespace a2m
