"""Host mirror of the reference's inference call sites (infer.py:37-44 predict_and_stitch,
audio_to_midi_dataset.py:277-294 window slicing), plus the batch partition used for multi-GPU runs."""
from __future__ import annotations

import math

import numpy as np

from . import modelutil
from .model import MODEL_AUDIO_LENGTH, SAMPLE_RATE, model_config, vmap
from .rope import precompute_frequencies


def slice_windows(audio_samples: np.ndarray, overlap: float = 0.25):
    """load_and_slice_full_audio without the ffmpeg decode: (2, N) -> ((W, 2, 80000), window seconds).
    `overlap` is in SECONDS (audio_to_midi_dataset.py:281-282)."""
    window = round(MODEL_AUDIO_LENGTH * SAMPLE_RATE)
    ov = round(overlap * SAMPLE_RATE)
    step = window - ov
    n = math.ceil((audio_samples.shape[1] - ov) / step)
    out = np.zeros((n, audio_samples.shape[0], window), dtype=np.float32)
    for i in range(n):
        seg = audio_samples[:, i * step:i * step + window]
        out[i, :, : seg.shape[1]] = seg
    return out, MODEL_AUDIO_LENGTH


def shard_windows(n_windows: int, world_size: int, rank: int):
    """Contiguous block of window indices for `rank` (weights replicated, no collective: SURVEY.md §8e).
    Blocks differ in size by at most one, earlier ranks take the extra window."""
    base, extra = divmod(n_windows, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def predict_and_stitch(model, state, samples, window_duration: float, overlap: float = 0.0, max_batch: int = 256):
    """infer.py:37-44: batched predict, fp32 probs, stitched probs, duration per frame."""
    rope_freqs = precompute_frequencies(model_config["attention_size"], 300)
    predict = vmap(model.predict, in_axes=(None, 0, None))
    chunks = []
    for i in range(0, samples.shape[0], max_batch):
        _logits, p = predict(state, samples[i:i + max_batch], rope_freqs)
        chunks.append(p.cpu().numpy() if hasattr(p, "cpu") else np.asarray(p))
    probs = np.concatenate(chunks).astype(np.float32)
    duration_per_frame = window_duration / probs.shape[1]
    return probs, modelutil.stitch_probs(probs, overlap, duration_per_frame), duration_per_frame


def prepare_windows_device(model, audio_samples, overlap: float = 0.25, device=None):
    """Device-side load_full_audio normalisation (python.rs:235-264) + load_and_slice_full_audio slicing
    (audio_to_midi_dataset.py:277-294): raw decoded clip (2, N) fp32 (numpy or torch CUDA) -> torch CUDA (W, 2, 80000)."""
    import ctypes as C
    import torch
    from . import _lib
    from .model import _default_device
    dev = _default_device() if device is None else device
    eng = model._engine(dev)
    tdev = torch.device(f"cuda:{dev}")
    clip = torch.as_tensor(np.ascontiguousarray(audio_samples, np.float32) if isinstance(audio_samples, np.ndarray) else audio_samples)
    clip = clip.to(tdev, torch.float32).contiguous()
    if clip.ndim != 2 or clip.shape[0] != 2:
        raise ValueError(f"audio must be (2, N), got {tuple(clip.shape)}")
    n = int(clip.shape[1])
    nw = int(eng.L.a2m_window_count(n, float(overlap)))
    if nw <= 0:
        raise ValueError("clip shorter than the overlap")
    out = torch.empty((nw, 2, 80000), dtype=torch.float32, device=tdev)
    stream = C.c_void_p(torch.cuda.current_stream(tdev).cuda_stream)
    rc = eng.L.a2m_prepare_windows(eng.h, clip.data_ptr(), n, float(overlap), out.data_ptr(), nw, stream)
    _lib.check(eng.h, rc, "a2m_prepare_windows")
    return out


def transcribe_clip(model, audio_samples, overlap: float = 0.25, max_batch: int = 64, rank: int = 0, world_size: int = 1):
    """Long-audio transcription (BASELINE config 5; infer.py:339 / audio_to_midi.py:38-53): normalise + slice on the
    device, batched forward of this rank's block of windows, then (rank 0 / single process) stitch and eventize.
    Returns (events, stitched_probs, probs_of_this_rank)."""
    import torch
    windows = prepare_windows_device(model, audio_samples, overlap)
    lo, hi = shard_windows(windows.shape[0], world_size, rank)
    rope_freqs = precompute_frequencies(model_config["attention_size"], 300)
    chunks = []
    for i in range(lo, hi, max_batch):
        _lg, p = model.predict(None, windows[i:min(i + max_batch, hi)], rope_freqs)
        chunks.append(p)
    probs = torch.cat(chunks).cpu().numpy().astype(np.float32) if chunks else np.zeros((0, 250, 90), np.float32)
    if world_size > 1:
        return None, None, probs          # the caller gathers the per-rank blocks in rank order, then stitches
    stitched = modelutil.stitch_probs(probs, overlap, MODEL_AUDIO_LENGTH / probs.shape[1])
    return modelutil.extract_events(stitched), stitched, probs
