"""Synthetic piano-like audio + labels of the model's window shape.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Deterministic: numpy PCG64(seed).
Audio is normalised exactly like the reference loader (python.rs:235-264: RMS over both
channels in f64, rounded through f16); labels follow python.rs:423-447 + the [0.005, 0.995]
clamp of label smoothing (python.rs:822-836, train.py:767).
"""
from __future__ import annotations

import numpy as np

from .events import normalize_audio, to_frame_events
from .params import SAMPLE_RATE, VOCAB, WINDOW_SAMPLES

FRAME_S = 0.02


def synth_notes(rng, duration_s=5.0, max_notes=6):
    n = int(rng.integers(1, max_notes + 1))
    notes = []
    for _ in range(n):
        key = int(rng.integers(21, 109))
        onset = float(rng.uniform(0.0, max(duration_s - 0.5, 0.1)))
        length = float(rng.uniform(0.2, 1.5))
        amp = float(rng.uniform(0.1, 1.0))
        notes.append((key, onset, length, amp))
    return notes


def render(notes, n_samples, rng, sample_rate=SAMPLE_RATE):
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    mono = np.zeros(n_samples, dtype=np.float64)
    for key, onset, _length, amp in notes:
        f0 = 440.0 * 2.0 ** ((key - 69) / 12.0)
        tt = t - onset
        env = np.where(tt >= 0.0, np.exp(-3.0 * np.clip(tt, 0.0, None)), 0.0)
        for h in range(1, 9):
            if f0 * h < sample_rate / 2:
                mono += amp / h * env * np.sin(2 * np.pi * f0 * h * tt)
    gl, gr = rng.uniform(0.5, 1.0, size=2)
    left = (gl * mono + rng.normal(0.0, 0.01, n_samples)).astype(np.float32)
    right = (gr * mono + rng.normal(0.0, 0.01, n_samples)).astype(np.float32)
    return left, right


def make_windows(batch: int, seed: int = 1234, with_labels: bool = False):
    """(B, 2, 80000) fp32 [, labels (B, 250, 90) fp32]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    audio = np.zeros((batch, 2, WINDOW_SAMPLES), dtype=np.float32)
    labels = np.zeros((batch, 250, VOCAB), dtype=np.float32) if with_labels else None
    for b in range(batch):
        notes = synth_notes(rng)
        left, right = render(notes, WINDOW_SAMPLES, rng)
        audio[b, 0], audio[b, 1] = normalize_audio(left, right)
        if with_labels:
            ev = [(int(round(on / FRAME_S)), key - 21, max(int(round(ln / FRAME_S)), 1), 7)
                  for key, on, ln, _ in notes]
            labels[b] = np.clip(to_frame_events(ev, 250), 0.005, 0.995)
    return (audio, labels) if with_labels else audio


def make_windows_fast(batch: int, seed: int = 1234):
    """Cheap stand-in with the same shape/statistics (unit RMS, f16-rounded) for large-batch benches:
    renders 8 distinct windows and tiles them with per-window gains, so values still differ per window."""
    base = make_windows(min(batch, 8), seed)
    if batch <= 8:
        return base
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    reps = -(-batch // base.shape[0])
    out = np.tile(base, (reps, 1, 1))[:batch].copy()
    gains = rng.uniform(0.8, 1.2, size=(batch, 1, 1)).astype(np.float32)
    return (out * gains).astype(np.float16).astype(np.float32)


def make_clip(duration_s: float, seed: int = 1234):
    """One long stereo clip (2, N) for the long-audio configuration (config 5)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = int(round(duration_s * SAMPLE_RATE))
    notes = []
    t0 = 0.0
    while t0 < duration_s:
        for key, onset, length, amp in synth_notes(rng):
            notes.append((key, t0 + onset, length, amp))
        t0 += 5.0
    # render in 5 s chunks with a 4 s tail to keep it cheap
    left = np.zeros(n, np.float32)
    right = np.zeros(n, np.float32)
    chunk = 5 * SAMPLE_RATE
    gl, gr = rng.uniform(0.5, 1.0, size=2)
    for c0 in range(0, n, chunk):
        c1 = min(c0 + chunk, n)
        t = np.arange(c0, c1, dtype=np.float64) / SAMPLE_RATE
        mono = np.zeros(c1 - c0)
        for key, onset, _l, amp in notes:
            if onset > t[-1] or onset < t[0] - 4.0:
                continue
            f0 = 440.0 * 2.0 ** ((key - 69) / 12.0)
            tt = t - onset
            env = np.where(tt >= 0.0, np.exp(-3.0 * np.clip(tt, 0.0, None)), 0.0)
            for h in range(1, 9):
                if f0 * h < SAMPLE_RATE / 2:
                    mono += amp / h * env * np.sin(2 * np.pi * f0 * h * tt)
        left[c0:c1] = gl * mono + rng.normal(0.0, 0.01, c1 - c0)
        right[c0:c1] = gr * mono + rng.normal(0.0, 0.01, c1 - c0)
    nl, nr = normalize_audio(left, right)
    return np.stack([nl, nr])
