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
