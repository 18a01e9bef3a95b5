// Host-side post-processing of the model's probabilities: the C++ home of what the reference keeps in its
// Rust `modelutil` crate (no Rust toolchain exists in this image).  Same semantics, new code:
//   a2m_stitch_probs     <- rust-plugins/src/common.rs:13-45
//   a2m_extract_events   <- rust-plugins/src/common.rs:47-144
//   a2m_to_frame_events  <- rust-plugins/src/python.rs:423-447 (+ :980-1005)
//   extract_midi_events / free_midi_events <- rust-plugins/src/cbinds.rs:51-91 (same symbols and struct layouts)
// These run on the host in the reference as well; they are not part of the CUDA hot path.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <tuple>
#include <vector>

#include "../../include/a2m.h"

namespace {

struct Ev {
  uint32_t attack, key, duration, velocity;
  bool operator<(const Ev& o) const {
    return std::tie(attack, key, duration, velocity) < std::tie(o.attack, o.key, o.duration, o.velocity);
  }
};

inline uint32_t span(int64_t end, int64_t start) { return static_cast<uint32_t>(std::max<int64_t>(end - start, 1)); }

// Highest probability reached from `frame` on while the curve keeps rising, tolerating dips for up to
// 10 frames (the activation look-ahead).
float peak_ahead(const float* p, int64_t frames, int64_t notes, int64_t frame, int64_t key) {
  float best = p[frame * notes + key];
  for (int64_t i = frame + 1; i < frames; ++i) {
    const float v = p[i * notes + key];
    if (v > best) best = v;
    else if (i - frame > 10) break;
  }
  return best;
}

std::vector<Ev> extract(const float* p, int64_t frames, int64_t notes) {
  const float kReGap = 0.1f, kReThresh = 0.4f, kOn = 0.5f, kOff = 0.1f;
  const uint32_t kVelocity = 7;  // the reference has not implemented velocity yet
  std::vector<Ev> out;
  std::vector<int64_t> started(static_cast<size_t>(notes), -1);
  for (int64_t f = 0; f < frames; ++f) {
    for (int64_t k = 0; k < notes; ++k) {
      const float cur = p[f * notes + k];
      const int64_t s = started[static_cast<size_t>(k)];
      if (s < 0) {
        if (cur > kOn) {
          started[static_cast<size_t>(k)] = f;
          (void)peak_ahead(p, frames, notes, f, k);  // feeds the (constant) velocity in the reference
        }
        continue;
      }
      if (cur < kOff) {  // released
        out.push_back({static_cast<uint32_t>(s), static_cast<uint32_t>(k), span(f, s), kVelocity});
        started[static_cast<size_t>(k)] = -1;
        continue;
      }
      // possible re-attack: mean of the next six frames exceeds the mean of the previous six by > gap.
      // Both sums are divided by six even when fewer than six frames remain.
      bool rising = false;
      if (static_cast<float>(f) - static_cast<float>(s) > 5.0f) {
        float before = 0.f, after = 0.f;
        for (int64_t i = f - 6; i < f; ++i) before += p[i * notes + k];
        before /= 6.0f;
        for (int64_t i = f; i < std::min<int64_t>(f + 6, frames); ++i) after += p[i * notes + k];
        after /= 6.0f;
        rising = (after - before) > kReGap;
      }
      if (f < frames - 1 && cur < p[(f + 1) * notes + k]) continue;  // decide on the local maximum
      if (cur > kReThresh && rising) {
        out.push_back({static_cast<uint32_t>(s), static_cast<uint32_t>(k), span(f - 1, s), kVelocity});
        started[static_cast<size_t>(k)] = f;
      }
    }
  }
  for (int64_t k = 0; k < notes; ++k) {
    const int64_t s = started[static_cast<size_t>(k)];
    if (s >= 0) out.push_back({static_cast<uint32_t>(s), static_cast<uint32_t>(k), span(frames, s), kVelocity});
  }
  std::sort(out.begin(), out.end());
  return out;
}

MidiEventList* to_list(const std::vector<Ev>& ev) {
  auto* list = static_cast<MidiEventList*>(std::malloc(sizeof(MidiEventList)));
  if (!list) return nullptr;
  list->length = ev.size();
  list->_capacity = ev.size();
  list->ptr = ev.empty() ? nullptr : static_cast<MidiEvent*>(std::calloc(ev.size(), sizeof(MidiEvent)));
  for (size_t i = 0; i < ev.size(); ++i) {
    list->ptr[i].attack_time = ev[i].attack;
    list->ptr[i].note = static_cast<uint8_t>(ev[i].key);
    list->ptr[i].duration = ev[i].duration;
    list->ptr[i].velocity = static_cast<uint8_t>(ev[i].velocity);
  }
  return list;
}

inline float half_to_float(uint16_t h) {
  const uint32_t sign = (h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {
      int e = -1;
      do { ++e; man <<= 1; } while ((man & 0x400u) == 0);
      bits = sign | static_cast<uint32_t>(127 - 15 - e) << 23 | (man & 0x3FFu) << 13;
    }
  } else if (exp == 31) bits = sign | 0x7F800000u | man << 13;
  else bits = sign | (exp + 112u) << 23 | man << 13;
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}

// Generic stitcher over an accessor so the fp32 and the strided-f16 entry points share one body.
template <class Get>
int64_t stitch(Get get, int64_t windows, int64_t frames, int64_t cats, double overlap, double dpf, float* out) {
  const double ov = overlap / dpf;  // overlapping frames, possibly fractional
  const int64_t out_frames = windows * frames - static_cast<int64_t>(ov) * (windows - 1);
  if (!out) return out_frames;
  std::fill(out, out + out_frames * cats, 0.0f);
  const int64_t blend_until = static_cast<int64_t>(std::ceil(ov));
  double base = 0.0;
  for (int64_t w = 0; w < windows; ++w) {
    const int64_t row0 = static_cast<int64_t>(base);
    for (int64_t f = 0; f < frames; ++f) {
      float* dst = out + (row0 + f) * cats;
      if (w > 0 && f <= blend_until) {
        const double t = static_cast<double>(f) / ov;  // 0/0 = NaN when overlap == 0, as in the reference
        for (int64_t c = 0; c < cats; ++c)
          dst[c] = static_cast<float>((1.0 - t) * static_cast<double>(dst[c]) + t * static_cast<double>(get(w, f, c)));
      } else {
        for (int64_t c = 0; c < cats; ++c) dst[c] = get(w, f, c);
      }
    }
    base += static_cast<double>(frames) - ov;
  }
  return out_frames;
}

}  // namespace

extern "C" {

int64_t a2m_stitch_probs(const float* probs, int64_t windows, int64_t frames, int64_t cats, double overlap,
                         double duration_per_frame, float* out) {
  // Same arithmetic as the generic stitcher above (common.rs:13-45), with the rows that are plain copies -- everything outside
  // the cross-fade of at most ceil(overlap) + 1 frames per window -- moved by memcpy: for a 10-minute clip that is 95 % of the
  // 2.7 M elements, and the element-wise path was the largest host cost of the long-clip configuration.
  const double ov = overlap / duration_per_frame;
  const int64_t out_frames = windows * frames - static_cast<int64_t>(ov) * (windows - 1);
  if (!out) return out_frames;
  std::fill(out, out + out_frames * cats, 0.0f);
  const int64_t blend_until = static_cast<int64_t>(std::ceil(ov));
  double base = 0.0;
  for (int64_t w = 0; w < windows; ++w) {
    const int64_t row0 = static_cast<int64_t>(base);
    const float* src = probs + w * frames * cats;
    int64_t f = 0;
    if (w > 0) {
      for (; f < frames && f <= blend_until; ++f) {
        float* dst = out + (row0 + f) * cats;
        const double t = static_cast<double>(f) / ov;  // 0/0 = NaN when overlap == 0, as in the reference
        for (int64_t c = 0; c < cats; ++c)
          dst[c] = static_cast<float>((1.0 - t) * static_cast<double>(dst[c]) + t * static_cast<double>(src[f * cats + c]));
      }
    }
    if (f < frames) std::memcpy(out + (row0 + f) * cats, src + f * cats, sizeof(float) * static_cast<size_t>((frames - f) * cats));
    base += static_cast<double>(frames) - ov;
  }
  return out_frames;
}

MidiEventList* a2m_extract_events(const float* probs, int64_t frames, int64_t notes) {
  return to_list(extract(probs, frames, notes));
}

MidiEventList* extract_midi_events(MLMultiArrayWrapper3 data, double overlap, double duration_per_frame) {
  const int64_t W = static_cast<int64_t>(data.dims[0]), F = static_cast<int64_t>(data.dims[1]),
                C = static_cast<int64_t>(data.dims[2]);
  const uint16_t* base = reinterpret_cast<const uint16_t*>(data.data);
  auto get = [&](int64_t w, int64_t f, int64_t c) {
    return half_to_float(base[w * data.strides[0] + f * data.strides[1] + c * data.strides[2]]);
  };
  const int64_t out_frames = stitch(get, W, F, C, overlap, duration_per_frame, nullptr);
  std::vector<float> st(static_cast<size_t>(std::max<int64_t>(out_frames, 0) * C));
  stitch(get, W, F, C, overlap, duration_per_frame, st.data());
  return to_list(extract(st.data(), out_frames, C));
}

void free_midi_events(MidiEventList* ptr) {
  if (!ptr) return;
  std::free(ptr->ptr);
  std::free(ptr);
}

int a2m_to_frame_events(const MidiEvent* events, int64_t n_events, int64_t frame_count, float* out) {
  if (!out || frame_count < 0) return A2M_EINVAL;
  const int64_t cats = A2M_VOCAB;
  std::fill(out, out + frame_count * cats, 0.0f);
  for (int64_t i = 0; i < n_events; ++i) {
    const int64_t start = static_cast<int64_t>(events[i].attack_time);
    const int64_t end = start + static_cast<int64_t>(events[i].duration);
    const int64_t key = events[i].note;
    if (key >= cats) return A2M_EINVAL;
    if (start > 0 && start < frame_count) out[(start - 1) * cats + key] = 0.0f;  // blank frame before an attack
    for (int64_t f = std::max<int64_t>(start, 0); f < std::min(end, frame_count); ++f) {
      const float t = static_cast<float>(f) - static_cast<float>(start);
      out[f * cats + key] = std::max(std::exp(-0.05f * t), 0.6f);
    }
  }
  return A2M_OK;
}

}  // extern "C"
