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
