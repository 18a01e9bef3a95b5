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
// F x notes elements in parallel and leaves three BIT MASKS per key, one bit per frame:
//     on   p > 0.5                                        (an idle key starts a note)
//     off  p < 0.1                                        (a sounding key is released)
//     re   p > 0.4, not p[f] < p[f+1], rise > 0.1         (a sounding key is re-attacked, if its note is older than 5 frames)
// Pass 2, one thread per key, is the machine itself -- the only sequential part -- but it no longer visits frames: an idle key
// jumps to the next set bit of `on`, a sounding key to the next set bit of `off | re` (find-first-set over 32 frames per word),
// so its cost is the number of events plus F / 32 words instead of F steps.  History: the comparisons inside one sequential walk
// per key, 16 ms for a 10-minute clip; a walk over precomputed flag bytes, 2.4 ms; 256-frame segments replayed from each key's
// latest p < 0.1 frame were SLOWER (5.5 ms) on tracks that rarely fall below 0.1, where every segment replays from frame 0.
constexpr int EX_KEYS = 96;                 // >= notes
constexpr int EVM_THREADS = 256;
constexpr uint32_t EVF_OFF = 1u, EVF_ON = 2u, EVF_RE = 4u;

// grid = ceil(F / 32) CTAs; CTA w owns frames 32 w .. 32 w + 31 of every key and writes word w of the three masks of every key:
// word (w & 3) of the 16-byte vector  masks4[(m * Q + (w >> 2)) * EX_KEYS + key],  m = 0 (on), 1 (off), 2 (re),  Q = ceil(F / 128):
// a vector holds 128 frames of one key, and the vectors of the keys are adjacent, so that the warp of pass 2 reads 512
// contiguous bytes per mask and step.  (The words past ceil(F / 32) are zeroed by the caller.)  Frames >= F contribute zero bits.
__global__ void __launch_bounds__(EVM_THREADS) event_masks_kernel(const float* __restrict__ probs, int F, int notes, int Q,
                                                                  uint32_t* __restrict__ masks) {
  __shared__ uint8_t tile[32][EX_KEYS + 4];
  const int w = blockIdx.x, f0 = w * 32;
  for (int i = threadIdx.x; i < 32 * notes; i += EVM_THREADS) {     // coalesced over the row-major [frame][key] array
    const int fr = i / notes, key = i - fr * notes, f = f0 + fr;
    uint32_t v = 0;
    if (f < F) {
      const float* p = probs + static_cast<long long>(f) * notes + key;    // p[k * notes] = this key, k frames later
      const float cur = p[0];
      if (cur < 0.1f) v |= EVF_OFF;
      if (cur > 0.5f) v |= EVF_ON;
      const bool defer = f < F - 1 && cur < p[notes];     // "handle the re-activation in the next frame where the probability is larger"
      if (cur > 0.4f && !defer && f >= 6) {
        float before = 0.f, after = 0.f;
        for (int k = -6; k < 0; ++k) before = __fadd_rn(before, p[static_cast<long long>(k) * notes]);
        before = __fdiv_rn(before, 6.0f);
        const int n = min(6, F - f);
        for (int k = 0; k < n; ++k) after = __fadd_rn(after, p[static_cast<long long>(k) * notes]);
        after = __fdiv_rn(after, 6.0f);
        if (__fsub_rn(after, before) > 0.1f) v |= EVF_RE;
      }
    }
    tile[fr][key] = static_cast<uint8_t>(v);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int key = warp; key < notes; key += EVM_THREADS / 32) {      // lane = frame inside the word
    const uint32_t v = tile[lane][key];
    const uint32_t on = __ballot_sync(0xffffffffu, v & EVF_ON), off = __ballot_sync(0xffffffffu, v & EVF_OFF),
                   re = __ballot_sync(0xffffffffu, v & EVF_RE);
    if (lane < 3) masks[((static_cast<size_t>(lane) * Q + (w >> 2)) * EX_KEYS + key) * 4 + (w & 3)] = lane == 0 ? on : (lane == 1 ? off : re);
  }
}

// One CTA, thread = key.  Each thread appends its key's events to its own list, staging[key * key_cap + i] (key_cap >= F / 2 + 2:
// a key closes at most one event every second frame); after a prefix sum over the keys the CTA copies the lists, one after the
// other and coalesced, to events[] as sortable 64-bit words  attack << 32 | key << 24 | duration  (frames < 2^24, checked by
// the caller), so that the host's ascending sort of the words IS the reference's (attack, key, duration) order (common.rs:142).
// *total = number of events; words beyond cap are dropped (the caller calls again with a larger buffer).
// The loop runs over the 128-frame vectors, the same one for every key, so the loads are coalesced, independent of the machine's
// state and fetched one vector ahead; only the frames a key reacts to inside a vector cost a (divergent) inner trip of ~25
// instructions on two 64-bit words.  Measured on a 10-minute clip whose keys hover around the thresholds (13 780 events): a scan
// loop nested inside a per-key event loop 1.2 ms (every outer trip lasted as long as the longest quiet stretch among the warp's
// 32 keys, each step a dependent L2 round trip).
__global__ void __launch_bounds__(EX_KEYS) extract_events_kernel(const uint4* __restrict__ masks4, int F, int notes, int Q,
                                                                 unsigned long long* __restrict__ staging, long long key_cap,
                                                                 unsigned long long* __restrict__ events, long long cap,
                                                                 int* __restrict__ total) {
  __shared__ int cnts[EX_KEYS];
  __shared__ long long offs[EX_KEYS + 1];
  const int key = threadIdx.x;
  const bool live = key < notes;
  const uint4* m_on = masks4 + key;
  const uint4* m_off = m_on + static_cast<size_t>(Q) * EX_KEYS;
  const uint4* m_re = m_off + static_cast<size_t>(Q) * EX_KEYS;
  unsigned long long* mine = staging + static_cast<size_t>(key) * key_cap;
  int count = 0;
  auto emit = [&](int start, int dur) {
    if (count < key_cap)
      mine[count] = (static_cast<unsigned long long>(start) << 32) | (static_cast<unsigned long long>(key) << 24) |
                    static_cast<unsigned long long>(dur);
    ++count;
  };
  auto u64 = [](uint32_t lo, uint32_t hi) { return static_cast<unsigned long long>(lo) | (static_cast<unsigned long long>(hi) << 32); };
  int started = -1;
  uint4 non = m_on[0], noff = m_off[0], nre = m_re[0];
  for (int q = 0; q < Q; ++q) {
    const uint4 von = non, voff = noff, vre = nre;
    if (q + 1 < Q) {
      const size_t o = static_cast<size_t>(q + 1) * EX_KEYS;
      non = m_on[o]; noff = m_off[o]; nre = m_re[o];
    }
    if (!live) continue;
    const unsigned long long on0 = u64(von.x, von.y), on1 = u64(von.z, von.w), off0 = u64(voff.x, voff.y), off1 = u64(voff.z, voff.w);
    const unsigned long long sr0 = off0 | u64(vre.x, vre.y), sr1 = off1 | u64(vre.z, vre.w);   // what a sounding key reacts to
    const int base = q << 7;
    unsigned long long keep0 = ~0ull, keep1 = ~0ull;        // frames of this vector the machine has not passed yet
    while (true) {
      const unsigned long long c0 = (started < 0 ? on0 : sr0) & keep0, c1 = (started < 0 ? on1 : sr1) & keep1;
      if ((c0 | c1) == 0ull) break;                         // nothing more in this vector (bits of frames >= F are never set)
      const int bit = c0 ? __ffsll(static_cast<long long>(c0)) - 1 : 64 + __ffsll(static_cast<long long>(c1)) - 1;
      const int g = base + bit;
      const bool is_off = (((bit < 64 ? off0 : off1) >> (bit & 63)) & 1ull) != 0ull;
      if (started < 0) {
        started = g;                                                             // p > 0.5: the note starts
      } else if (is_off) {                                                       // released (common.rs:81-84); tested first, as the reference does
        emit(started, max(g - started, 1));
        started = -1;
      } else if (g - started > 5) {                                              // re-attack (common.rs:88-123)
        emit(started, max(g - 1 - started, 1));
        started = g;
      }
      // the machine moves on to frame g + 1
      if (bit < 64) {
        keep0 = bit == 63 ? 0ull : (~0ull << (bit + 1));
      } else {
        keep0 = 0ull;
        keep1 = bit == 127 ? 0ull : (~0ull << (bit - 63));
      }
    }
  }
  if (live && started >= 0) emit(started, max(F - started, 1));                  // still sounding at the end of the track
  cnts[key] = live ? count : 0;
  __syncthreads();
  if (key == 0) {
    long long run = 0;
    for (int k = 0; k < EX_KEYS; ++k) { offs[k] = run; run += cnts[k]; }
    offs[EX_KEYS] = run;
    *total = static_cast<int>(run < 0x7fffffffll ? run : 0x7fffffffll);
  }
  __syncthreads();
  for (int k = 0; k < notes; ++k) {
    const unsigned long long* src = staging + static_cast<size_t>(k) * key_cap;
    const long long o = offs[k];
    for (int i = key; i < cnts[k]; i += EX_KEYS)
      if (o + i < cap) events[o + i] = src[i];
  }
}

}  // namespace a2m
