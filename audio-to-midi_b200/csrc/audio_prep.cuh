// Step before the hot path, on the device: loudness normalisation of a whole clip
// (rust-plugins/src/python.rs:235-264: peak test, 1 / sqrt(mean square over both channels) in f64, round to f16) fused
// with the window slicing of load_and_slice_full_audio (audio_to_midi_dataset.py:277-294: windows of 80 000 samples every
// 80 000 - round(overlap_s * 16 000) samples, the last one zero padded).  Two launches per clip: a statistics pass
// (peak, sum of squares in f64) and a pass that writes the [W, 2, 80000] fp32 window tensor the model forward reads, so
// a 10-minute clip (9.6 M samples per channel) never takes a host pass.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace a2m {

struct ClipStats {
  double sumsq;          // sum over both channels of x^2
  unsigned int peak;     // bit pattern of max |x| (non-negative floats order like unsigned ints)
  unsigned int pad;
};

__global__ void __launch_bounds__(256) clip_stats_kernel(const float* __restrict__ clip, long long n_total, ClipStats* __restrict__ st) {
  double s = 0.0;
  float mx = 0.f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n_total; i += static_cast<long long>(gridDim.x) * 256ll) {
    const float v = clip[i];
    s += static_cast<double>(v) * static_cast<double>(v);
    mx = fmaxf(mx, fabsf(v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ double ss[8];
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sm[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    float m = 0.f;
    for (int i = 0; i < 8; ++i) { t += ss[i]; m = fmaxf(m, sm[i]); }
    atomicAdd(&st->sumsq, t);
    atomicMax(&st->peak, __float_as_uint(m));
  }
}

// windows[w][c][i] = f16(clip[c][w * step + i] * adj) widened to fp32, 0 beyond the end of the clip
__global__ void __launch_bounds__(256) slice_normalize_kernel(const float* __restrict__ clip, long long n_samples, int step, int window,
                                                              int n_windows, const ClipStats* __restrict__ st, float* __restrict__ out) {
  const float peak = __uint_as_float(st->peak);
  const bool scale = peak > 0.05f;
  const double adj = scale ? sqrt(1.0 / (st->sumsq / (2.0 * static_cast<double>(n_samples)))) : 1.0;
  const long long total = static_cast<long long>(n_windows) * 2 * window;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256ll) {
    const int s = static_cast<int>(i % window);
    const int c = static_cast<int>((i / window) & 1);
    const long long w = i / (2ll * window);
    const long long src = w * step + s;
    float v = 0.f;
    if (src < n_samples) {
      const float x = clip[static_cast<long long>(c) * n_samples + src];
      v = scale ? __half2float(__double2half(static_cast<double>(x) * adj)) : __half2float(__float2half_rn(x));
    }
    out[i] = v;
  }
}

}  // namespace a2m
