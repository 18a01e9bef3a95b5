// Validation metrics on the device: infer.py:94-158 detailed_event_loss for a whole batch of windows in one
// launch.  The reference does it serially on the host, one window at a time: modelutil.extract_events (common.rs:47-144) ->
// modelutil.to_frame_events (python.rs:423-447, 980-1005) -> |predicted - expected| sums.
//
// extract_events is independent per key: each of the 90 keys of a window is a hysteresis state machine over that key's frames
// (thresholds 0.5 on / 0.1 off / 0.4 + rising-mean gap 0.1 for a re-attack decided on the local maximum).  One thread owns one
// (window, key) column: it runs the state machine in the reference's f32 arithmetic (sequential sums, IEEE division), rasterises
// every event it emits into its column of `pred` as to_frame_events does, and accumulates the comparison against the annotation.
// A CTA is one window; thread 0 adds the 90 per-key partial sums in key order (double), so the result does not depend on scheduling.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace a2m {

constexpr int EM_THREADS = 96;     // >= A2M_VOCAB keys
constexpr int EM_DECAY = 11;       // exp(-0.05 t) > 0.6 only for t <= 10 (python.rs:441-444: max(exp(-0.05 t), 0.6))

struct EventDecay {
  float v[EM_DECAY];               // correctly rounded f32 exp(-0.05f * t), computed on the host
};

// probs, expected, pred: [B, F, notes] fp32; metrics: [B, 5] = full_diff, phantom_notes_diff, missed_notes_diff, notes_hit, hit_rate;
// n_events (optional): [B, notes] events emitted per key.
__global__ void __launch_bounds__(EM_THREADS) event_metrics_kernel(const float* __restrict__ probs, const float* __restrict__ expected,
                                                                   int F, int notes, float* __restrict__ pred, float* __restrict__ metrics,
                                                                   int* __restrict__ n_events, const EventDecay decay) {
  const int b = blockIdx.x, key = threadIdx.x;
  __shared__ double part[EM_THREADS][4];
  double full = 0.0, phantom = 0.0, missed = 0.0, hit = 0.0;
  if (key < notes) {
    const float* p = probs + static_cast<size_t>(b) * F * notes + key;
    float* q = pred + static_cast<size_t>(b) * F * notes + key;
    for (int f = 0; f < F; ++f) q[static_cast<size_t>(f) * notes] = 0.f;
    int count = 0;
    auto emit = [&](int start, int dur) {                       // convert_to_frame_events, python.rs:423-447
      if (start > 0 && start < F) q[static_cast<size_t>(start - 1) * notes] = 0.f;
      const int end = min(start + dur, F);
      for (int f = start; f < end; ++f) {
        const int t = f - start;
        q[static_cast<size_t>(f) * notes] = t < EM_DECAY ? fmaxf(decay.v[t], 0.6f) : 0.6f;
      }
      ++count;
    };
    int started = -1;
    for (int f = 0; f < F; ++f) {
      const float cur = p[static_cast<size_t>(f) * notes];
      if (started < 0) {
        if (cur > 0.5f) started = f;                            // the look-ahead peak only feeds the (constant) velocity
        continue;
      }
      if (cur < 0.1f) {                                         // released (common.rs:81-84)
        emit(started, max(f - started, 1));
        started = -1;
        continue;
      }
      bool rising = false;
      if (static_cast<float>(f) - static_cast<float>(started) > 5.0f) {   // common.rs:93-112; both means divide by six
        float before = 0.f, after = 0.f;
        for (int i = f - 6; i < f; ++i) before = __fadd_rn(before, p[static_cast<size_t>(i) * notes]);
        before = __fdiv_rn(before, 6.0f);
        const int hi = min(f + 6, F);
        for (int i = f; i < hi; ++i) after = __fadd_rn(after, p[static_cast<size_t>(i) * notes]);
        after = __fdiv_rn(after, 6.0f);
        rising = __fsub_rn(after, before) > 0.1f;
      }
      if (f < F - 1 && cur < p[static_cast<size_t>(f + 1) * notes]) continue;   // decided on the local maximum (common.rs:114-117)
      if (cur > 0.4f && rising) {
        emit(started, max(f - 1 - started, 1));
        started = f;
      }
    }
    if (started >= 0) emit(started, max(F - started, 1));
    if (n_events) n_events[b * notes + key] = count;
    // detailed_event_loss, infer.py:111-130
    const float* e = expected + static_cast<size_t>(b) * F * notes + key;
    for (int f = 0; f < F; ++f) {
      const float pr = q[static_cast<size_t>(f) * notes], ex = e[static_cast<size_t>(f) * notes];
      full += static_cast<double>(fabsf(pr - ex));
      const bool pp = pr > 0.f, pe = ex > 0.f;
      if (pp && !pe) phantom += 1.0;
      if (pe && !pp) missed += static_cast<double>(ex);
      if (pp && pe) hit += 1.0;
    }
  }
  part[key][0] = full; part[key][1] = phantom; part[key][2] = missed; part[key][3] = hit;
  __syncthreads();
  if (key == 0) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < notes; ++k)
      for (int j = 0; j < 4; ++j) s[j] += part[k][j];
    float* m = metrics + static_cast<size_t>(b) * 5;
    m[0] = static_cast<float>(s[0]); m[1] = static_cast<float>(s[1]); m[2] = static_cast<float>(s[2]); m[3] = static_cast<float>(s[3]);
    const double denom = s[3] + s[1] + s[2];
    m[4] = denom > 0.0 ? static_cast<float>(s[3] / denom) : 1.0f;
  }
}

}  // namespace a2m

// ============================================================================================= long-clip post-processing
// stitch_probs (common.rs:13-45) and extract_events (common.rs:47-144) on the device, so that the probabilities of a long clip
// (config 5: 134 windows -> 30 175 stitched frames) never travel to the host: only the event list does.
namespace a2m {

// out[row, c] for every stitched row.  The reference writes windows in order into a zero-filled track: window w, frame f goes to
// row row0[w] + f (row0 = the truncated running sum of frames - overlap, accumulated in f64 on the host exactly as common.rs:41
// does); for w > 0 and f <= ceil(overlap) the write is the f64 blend (1 - f / ov) * old + (f / ov) * new.  Every row is therefore
// decided by the LAST window that covers it, and `old` is the previous window's plain value (or 0 where that window has ended);
// the host only takes this path when cross-fades cannot chain (frames - ceil(ov) - 1 > ceil(ov)).  No FMA contraction: the host
// code rounds the two products and the sum separately.
__global__ void __launch_bounds__(256) stitch_probs_kernel(const float* __restrict__ probs, const long long* __restrict__ row0, int W, int F,
                                                           int cats, double ov, int blend_until, long long out_frames,
                                                           float* __restrict__ out) {
  const long long total = out_frames * cats;
  const double step = static_cast<double>(F) - ov;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const long long r = i / cats;
    const int c = static_cast<int>(i - r * cats);
    int w = step > 0.0 ? static_cast<int>(static_cast<double>(r) / step) : 0;
    w = min(max(w, 0), W - 1);
    while (w + 1 < W && row0[w + 1] <= r) ++w;
    while (w > 0 && row0[w] > r) --w;
    const long long f = r - row0[w];
    float v = 0.f;
    if (f < F) {
      const float x = probs[(static_cast<long long>(w) * F + f) * cats + c];
      if (w > 0 && f <= blend_until) {
        const long long fp = r - row0[w - 1];
        const float old = fp < F ? probs[(static_cast<long long>(w - 1) * F + fp) * cats + c] : 0.f;
        const double t = static_cast<double>(f) / ov;       // 0 / 0 = NaN when overlap == 0, as in the reference
        v = static_cast<float>(__dadd_rn(__dmul_rn(1.0 - t, static_cast<double>(old)), __dmul_rn(t, static_cast<double>(x))));
      } else {
        v = x;
      }
    }
    out[i] = v;
  }
}

// ---- extract_events in two passes -----------------------------------------------------------------------------------------
// Every comparison the state machine makes is a pure function of the probabilities around a frame, not of the machine's state
// (common.rs:81-119): p < 0.1, p > 0.5, p > 0.4, p[f] < p[f+1], and the re-attack rise mean(p[f..f+6)) - mean(p[f-6..f)) > 0.1
// (both sums divided by six, sequential f32 adds starting from 0, as the reference forms them).  Pass 1 evaluates them for all
// F x notes elements in parallel into one flag byte each; pass 2, one thread per key, walks its key's bytes -- the only sequential
// part -- with a handful of integer instructions per frame.  (A first version evaluated the comparisons inside the sequential walk:
// 16 ms for a 10-minute clip, almost all of it exposed latency of three lonely warps.)
constexpr uint32_t EVF_OFF = 1u, EVF_ON = 2u, EVF_RE = 4u, EVF_DEFER = 8u, EVF_RISE = 16u;

__global__ void __launch_bounds__(256) event_flags_kernel(const float* __restrict__ probs, int F, int notes, uint8_t* __restrict__ flags) {
  const long long total = static_cast<long long>(F) * notes;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int f = static_cast<int>(i / notes);
    const float* p = probs + i;                       // p[k * notes] = this key, k frames later
    const float cur = p[0];
    uint32_t v = 0;
    if (cur < 0.1f) v |= EVF_OFF;
    if (cur > 0.5f) v |= EVF_ON;
    if (cur > 0.4f) v |= EVF_RE;
    if (f < F - 1 && cur < p[notes]) v |= EVF_DEFER;  // "handle the re-activation in the next frame where the probability is larger"
    if (f >= 6) {
      float before = 0.f, after = 0.f;
      for (int k = -6; k < 0; ++k) before = __fadd_rn(before, p[static_cast<long long>(k) * notes]);
      before = __fdiv_rn(before, 6.0f);
      const int n = min(6, F - f);
      for (int k = 0; k < n; ++k) after = __fadd_rn(after, p[static_cast<long long>(k) * notes]);
      after = __fdiv_rn(after, 6.0f);
      if (__fsub_rn(after, before) > 0.1f) v |= EVF_RISE;
    }
    flags[i] = static_cast<uint8_t>(v);
  }
}

constexpr int EX_TILE = 256;     // frames per shared-memory tile of flag bytes
constexpr int EX_KEYS = 96;      // >= notes
constexpr int EX_THREADS = 256;
constexpr int EX_WORDS = EX_TILE * EX_KEYS / 4;                          // tile capacity in 32-bit words
constexpr int EX_LD = (EX_WORDS + EX_THREADS - 1) / EX_THREADS;

// One CTA: threads 0 .. notes-1 each walk the flag bytes of their key; all 256 threads move the flags through a double-buffered
// shared-memory tile (a tile is one contiguous range of the [F, notes] byte array: flat, coalesced word copies whose loads are in
// flight while the walk runs).  `flags` must be readable up to the next multiple of 4 bytes past F * notes.
// events[key][k] = (attack, duration) for k < counts[key] <= cap; counts may exceed cap (overflow: the caller re-runs on the host).
__global__ void __launch_bounds__(EX_THREADS) extract_events_kernel(const uint8_t* __restrict__ flags, int F, int notes, uint2* __restrict__ events,
                                                                    int* __restrict__ counts, int cap) {
  __shared__ uint32_t tile[2][EX_WORDS];
  const int tid = threadIdx.x, key = threadIdx.x;
  const int tile_bytes = EX_TILE * notes;                                // multiple of 4: EX_TILE is
  const int tile_words = tile_bytes / 4;
  const long long total_words = (static_cast<long long>(F) * notes + 3) / 4;
  const uint32_t* gw = reinterpret_cast<const uint32_t*>(flags);
  uint32_t reg[EX_LD];
  auto gload = [&](int f0) {
    const long long w0 = static_cast<long long>(f0) * notes / 4;
#pragma unroll
    for (int k = 0; k < EX_LD; ++k) {
      const int idx = tid + k * EX_THREADS;
      reg[k] = (idx < tile_words && w0 + idx < total_words) ? gw[w0 + idx] : 0u;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int k = 0; k < EX_LD; ++k) {
      const int idx = tid + k * EX_THREADS;
      if (idx < tile_words) tile[buf][idx] = reg[k];
    }
  };
  int started = -1, count = 0;
  auto emit = [&](int start, int dur) {
    if (count < cap) events[static_cast<size_t>(key) * cap + count] = make_uint2(static_cast<uint32_t>(start), static_cast<uint32_t>(dur));
    ++count;
  };
  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int f0 = 0; f0 < F; f0 += EX_TILE) {
    const bool more = f0 + EX_TILE < F;
    if (more) gload(f0 + EX_TILE);
    if (key < notes) {
      const uint8_t* t = reinterpret_cast<const uint8_t*>(tile[buf]) + key;
      const int fend = min(f0 + EX_TILE, F);
      for (int f = f0; f < fend; ++f) {
        const uint32_t v = t[(f - f0) * notes];
        if (started < 0) {
          if (v & EVF_ON) started = f;
        } else if (v & EVF_OFF) {                                        // released (common.rs:81-84)
          emit(started, max(f - started, 1));
          started = -1;
        } else if (!(v & EVF_DEFER) && (v & EVF_RE) && (v & EVF_RISE) && f - started > 5) {   // re-attack (common.rs:88-123)
          emit(started, max(f - 1 - started, 1));
          started = f;
        }
      }
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  if (key < notes) {
    if (started >= 0) emit(started, max(F - started, 1));
    counts[key] = count;
  }
}

}  // namespace a2m
