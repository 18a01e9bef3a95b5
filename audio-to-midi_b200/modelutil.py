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


_EVENT_DTYPE = np.dtype({"names": ["attack_time", "note", "duration", "velocity"], "formats": ["<u8", "u1", "<u8", "u1"],
                         "offsets": [0, 8, 16, 24], "itemsize": 32})
assert C.sizeof(MidiEvent) == _EVENT_DTYPE.itemsize


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
        n = int(lst.contents.length)
        if n == 0:
            return []
        # one vectorised view of the #[repr(C)] MidiEvent array (cbinds.rs:9-15: u64, u8, u64, u8 at offsets 0, 8, 16, 24 of 32
        # bytes) instead of four ctypes attribute reads per event: a 10-minute clip yields > 10 000 events
        buf = (C.c_char * (n * C.sizeof(MidiEvent))).from_address(C.addressof(lst.contents.ptr.contents))
        ev = np.frombuffer(buf, dtype=_EVENT_DTYPE, count=n)
        return list(zip(ev["attack_time"].tolist(), ev["note"].tolist(), ev["duration"].tolist(), ev["velocity"].tolist()))
    finally:
        L.free_midi_events(lst)


def to_frame_events(all_events, frame_count: int) -> list:
    L = lib()
    out = []
    for events in all_events:
        rec = np.zeros(max(len(events), 1), dtype=_EVENT_DTYPE)
        if len(events):
            cols = np.asarray(events, dtype=np.int64).reshape(len(events), 4)
            rec["attack_time"], rec["note"], rec["duration"], rec["velocity"] = cols[:, 0], cols[:, 1], cols[:, 2], cols[:, 3]
        arr = rec.ctypes.data_as(C.POINTER(MidiEvent))
        frames = np.empty((frame_count, 90), dtype=np.float32)
        rc = L.a2m_to_frame_events(arr, len(events), frame_count, frames.ctypes.data)
        if rc != 0:
            raise ValueError("a2m_to_frame_events: bad event (key >= 90?)")
        out.append(frames)
    return out
