"""Restatement of the reference's Rust post-processing (the consumers of the hot path's output).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: `modelutil` is Rust
(ndarray 0.15.6, half 2.4.1, num-traits 0.2.18 per rust-plugins/Cargo.lock) and there is no
cargo/rustc here; the arithmetic is plain f32/f64 and fully visible in the cited lines.

  stitch_probs            rust-plugins/src/common.rs:13-45
  extract_events          rust-plugins/src/common.rs:47-144
  convert_to_frame_events rust-plugins/src/python.rs:423-447 (to_frame_events: python.rs:980-1005)
  normalize_audio         rust-plugins/src/python.rs:235-264
  slice_windows           audio_to_midi_dataset.py:277-294
  detailed_event_loss     infer.py:94-158 (the numeric fields only)

Pure-Python loops: use on small inputs (thousands of frames), not whole datasets.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def stitch_probs(all_probs: np.ndarray, overlap: float, duration_per_frame: float) -> np.ndarray:
    num_windows, frames_per_window, cats = all_probs.shape
    overlapping = float(overlap) / float(duration_per_frame)
    out_frames = num_windows * frames_per_window - int(overlapping) * (num_windows - 1)
    stitched = np.zeros((out_frames, cats), dtype=np.float32)
    base = 0.0
    ceil_ov = math.ceil(overlapping)
    with np.errstate(invalid="ignore", divide="ignore"):
        for w in range(num_windows):
            for frame in range(frames_per_window):
                row = int(base) + frame
                if w > 0 and frame <= ceil_ov:
                    # frame / overlapping is 0/0 = NaN when overlap == 0 (kept: the reference does it)
                    blend = np.float64(frame) / np.float64(overlapping)
                    old = stitched[row].astype(np.float64)
                    new = all_probs[w, frame].astype(np.float32).astype(np.float64)
                    stitched[row] = ((1.0 - blend) * old + blend * new).astype(np.float32)
                else:
                    stitched[row] = all_probs[w, frame].astype(np.float32)
            base += float(frames_per_window) - overlapping
    return stitched


def extract_events(probs: np.ndarray) -> list[tuple[int, int, int, int]]:
    probs = np.asarray(probs, dtype=np.float32)
    reactivation_gap, reactivation_threshold = f32(0.1), f32(0.4)
    activation_threshold, deactivation_threshold = f32(0.5), f32(0.1)
    num_frames, num_notes = probs.shape
    events = []

    def duration(end, start):
        return max(int(end) - int(start), 1)

    def activation_prob(frame, key):
        a = probs[frame, key]
        for i in range(frame + 1, num_frames):
            if probs[i, key] > a:
                a = probs[i, key]
            elif i - frame > 10:
                break
        return a

    playing = [None] * num_notes
    for frame in range(num_frames):
        for key in range(num_notes):
            p = probs[frame, key]
            cur = playing[key]
            if cur is not None:
                started_at, act = cur
                if p < deactivation_threshold:
                    events.append((started_at, key, duration(frame, started_at), 7))
                    playing[key] = None
                else:
                    since = f32(frame) - f32(started_at)
                    should = False
                    if since > 5.0:
                        prev = f32(0.0)
                        for i in range(frame - 6, frame):
                            prev = f32(prev + probs[i, key])
                        prev = f32(prev / f32(6.0))
                        nxt = f32(0.0)
                        for i in range(frame, min(frame + 6, num_frames)):
                            nxt = f32(nxt + probs[i, key])
                        nxt = f32(nxt / f32(6.0))
                        should = f32(nxt - prev) > reactivation_gap
                    if frame < num_frames - 1 and p < probs[frame + 1, key]:
                        continue
                    if p > reactivation_threshold and should:
                        events.append((started_at, key, duration(frame - 1, started_at), 7))
                        playing[key] = (frame, activation_prob(frame, key))
            elif p > activation_threshold:
                playing[key] = (frame, activation_prob(frame, key))
    for key in range(num_notes):
        if playing[key] is not None:
            started_at, _ = playing[key]
            events.append((started_at, key, duration(num_frames, started_at), 7))
    events.sort()
    return [(int(a), int(b), int(c), int(d)) for a, b, c, d in events]


def to_frame_events(events, frame_count: int, num_event_types: int = 90) -> np.ndarray:
    """convert_to_frame_events(events, frame_count, start_frame=0, num_frames_with_backing_samples=frame_count)."""
    frames = np.zeros((frame_count, num_event_types), dtype=np.float32)
    for attack, key, dur, _vel in events:
        start = int(attack)
        end = start + int(dur)
        if 0 < start < frame_count:
            frames[start - 1, key] = 0.0
        for fr in range(max(start, 0), min(end, frame_count)):
            t = f32(fr) - f32(start)
            # f32 exp as a correctly rounded libm expf gives it (Rust f32::exp -> expf): evaluate in f64, round once
            frames[fr, key] = max(f32(np.exp(np.float64(f32(-0.05) * t))), f32(0.6))
    return frames


def normalize_audio(left: np.ndarray, right: np.ndarray):
    """python.rs:235-264: RMS-normalise over both channels in f64, round to f16 (returned widened to f32)."""
    left = np.asarray(left, np.float32)
    right = np.asarray(right, np.float32)
    total_max = max(np.abs(left).max(), np.abs(right).max())
    if total_max <= 0.05:
        return left.astype(np.float16).astype(np.float32), right.astype(np.float16).astype(np.float32)
    n = float(left.size + right.size)
    variance = float(np.sum((left.astype(np.float64) ** 2 + right.astype(np.float64) ** 2) / n))
    adj = math.sqrt(1.0 / variance)
    nl = (left.astype(np.float64) * adj).astype(np.float16).astype(np.float32)
    nr = (right.astype(np.float64) * adj).astype(np.float16).astype(np.float32)
    return nl, nr


def slice_windows(audio: np.ndarray, overlap: float = 0.25, sample_rate: int = 16000, window_s: float = 5.0):
    """audio (2, N) -> (W, 2, 80000); ``overlap`` in seconds; last window zero padded."""
    window = round(window_s * sample_rate)
    ov = round(overlap * sample_rate)
    step = window - ov
    n_windows = math.ceil((audio.shape[1] - ov) / step)
    out = np.zeros((n_windows, audio.shape[0], window), dtype=audio.dtype)
    for i in range(n_windows):
        seg = audio[:, i * step:i * step + window]
        out[i, :, : seg.shape[1]] = seg
    return out


def detailed_event_loss(output_probs: np.ndarray, expected: np.ndarray) -> dict:
    predicted = to_frame_events(extract_events(output_probs), output_probs.shape[0])
    expected = expected[: predicted.shape[0]]
    full_diff = float(np.sum(np.abs(predicted - expected)))
    pp, pe = predicted > 0, expected > 0
    phantom = float(np.sum(pp & ~pe))
    missed = float(np.sum(expected[pe & ~pp]))
    hit = float(np.sum(pp & pe))
    denom = hit + phantom + missed                      # infer.py:127-130
    return {"full_diff": full_diff, "phantom_notes_diff": phantom, "missed_notes_diff": missed,
            "notes_hit": hit, "hit_rate": hit / denom if denom > 0 else 1.0}
