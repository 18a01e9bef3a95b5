"""Python face of the C++ post-processing (csrc/modelutil.cpp), with the call shapes of the reference's
PyO3 module `modelutil` (rust-plugins/src/python.rs:962-1005) so infer.py-style callers keep working:

    stitch_probs(f32[W, F, 90], overlap_s, duration_per_frame) -> f32[F', 90]
    extract_events(f32[F, 90]) -> list[(attack, key, duration, velocity)]
    to_frame_events(list[list[tuple]], frame_count) -> list[f32[F, 90]]
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import MidiEvent, lib


def stitch_probs(all_probs, overlap: float, duration_per_frame: float) -> np.ndarray:
    p = np.ascontiguousarray(all_probs, dtype=np.float32)
    if p.ndim != 3:
        raise ValueError("stitch_probs expects (windows, frames, categories)")
    w, f, c = p.shape
    L = lib()
    n = L.a2m_stitch_probs(p.ctypes.data, w, f, c, float(overlap), float(duration_per_frame), None)
    out = np.empty((n, c), dtype=np.float32)
    L.a2m_stitch_probs(p.ctypes.data, w, f, c, float(overlap), float(duration_per_frame), out.ctypes.data)
    return out


def extract_events(probs) -> list:
    p = np.ascontiguousarray(probs, dtype=np.float32)
    if p.ndim != 2:
        raise ValueError("extract_events expects (frames, notes)")
    L = lib()
    lst = L.a2m_extract_events(p.ctypes.data, p.shape[0], p.shape[1])
    if not lst:
        raise MemoryError("a2m_extract_events returned NULL")
    try:
        n = lst.contents.length
        ev = lst.contents.ptr
        return [(int(ev[i].attack_time), int(ev[i].note), int(ev[i].duration), int(ev[i].velocity)) for i in range(n)]
    finally:
        L.free_midi_events(lst)


def to_frame_events(all_events, frame_count: int) -> list:
    L = lib()
    out = []
    for events in all_events:
        arr = (MidiEvent * max(len(events), 1))()
        for i, (a, k, d, v) in enumerate(events):
            arr[i].attack_time, arr[i].note, arr[i].duration, arr[i].velocity = int(a), int(k), int(d), int(v)
        frames = np.empty((frame_count, 90), dtype=np.float32)
        rc = L.a2m_to_frame_events(arr, len(events), frame_count, frames.ctypes.data)
        if rc != 0:
            raise ValueError("a2m_to_frame_events: bad event (key >= 90?)")
        out.append(frames)
    return out
