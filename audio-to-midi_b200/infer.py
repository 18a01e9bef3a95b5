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
    chunks = []
    if isinstance(samples, np.ndarray) and samples.shape[0] > max_batch:
        # host windows in several batches: two batches in flight (H2D / D2H of one under the kernels of the other) instead of
        # one synchronous call per batch -- 4x the throughput of the loop below at 64 windows per batch (bench.py e2e)
        parts = (samples[i:i + max_batch] for i in range(0, samples.shape[0], max_batch))
        for _logits, p in model.predict_pipelined(parts, rope_freqs, state=state, copy=True):
            chunks.append(p)
    else:
        predict = vmap(model.predict, in_axes=(None, 0, None))
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


def detailed_event_loss(output_probs: np.ndarray, expected: np.ndarray) -> dict:
    """infer.py:94-158 without the plot: eventize the probabilities, rasterise them back to frames and compare with the
    annotation: full_diff, phantom / missed note mass, notes hit, hit_rate = hit / (hit + phantom + missed)."""
    output_probs = np.ascontiguousarray(output_probs, np.float32)
    predicted = modelutil.to_frame_events([modelutil.extract_events(output_probs)], output_probs.shape[0])[0]
    expected = np.asarray(expected)[: predicted.shape[0]]
    pp, pe = predicted > 0, expected > 0
    phantom = float(np.sum(pp & ~pe))
    missed = float(np.sum(expected[pe & ~pp]))
    hit = float(np.sum(pp & pe))
    denom = hit + phantom + missed
    return {"full_diff": float(np.sum(np.abs(predicted - expected))), "phantom_notes_diff": phantom,
            "missed_notes_diff": missed, "notes_hit": hit, "hit_rate": hit / denom if denom > 0 else 1.0}


def compute_testset_loss(model, audio, events, rank: int = 0, world_size: int = 1, max_batch: int = 64):
    """Validation pass of config 3 (compute_testset_loss_individual, train.py:86-209; infer.py:94-158): the annotated
    windows are batch-partitioned over ranks (contiguous blocks, no collective); each rank runs the batched forward on
    its block, the per-window BCE sum on the device (a2m_window_losses) and the event metrics on the host.
    audio (N, 2, 80000), events (N, 250, 90), numpy.  Returns (lo, hi, losses[hi-lo], [detailed_event_loss dict])."""
    import ctypes as C
    import torch
    from . import _lib
    from .model import _default_device
    lo, hi = shard_windows(audio.shape[0], world_size, rank)
    dev = _default_device()
    eng = model._engine(dev)
    tdev = torch.device(f"cuda:{dev}")
    rope_freqs = precompute_frequencies(model_config["attention_size"], 300)
    losses, details = [], []
    for i in range(lo, hi, max_batch):
        j = min(i + max_batch, hi)
        x = torch.as_tensor(np.ascontiguousarray(audio[i:j], np.float32)).to(tdev)
        y = torch.as_tensor(np.ascontiguousarray(events[i:j], np.float32)).to(tdev)
        logits, probs = model.predict(None, x, rope_freqs)
        out = torch.empty(j - i, dtype=torch.float32, device=tdev)
        stream = C.c_void_p(torch.cuda.current_stream(tdev).cuda_stream)
        _lib.check(eng.h, eng.L.a2m_window_losses(eng.h, logits.data_ptr(), y.data_ptr(), j - i, out.data_ptr(), stream), "a2m_window_losses")
        losses.append(out.cpu().numpy())
        pr = probs.cpu().numpy()
        details.extend(detailed_event_loss(pr[k], events[i + k]) for k in range(j - i))
    return lo, hi, (np.concatenate(losses) if losses else np.zeros(0, np.float32)), details


# ------------------------------------------------------------------------------------------ MIDI writer
NUM_VELOCITY_CATEGORIES = 10      # audio_to_midi_dataset.py:33
_TICKS_PER_BEAT = 480             # mido.MidiFile default
_TEMPO_US = 500000                # mido.bpm2tempo(120), 4/4  (infer.py:53-57)


def _varlen(n: int) -> bytes:
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


def write_midi_file(events, duration_per_frame: float, output_file: str):
    """infer.py:46-83 without mido (absent from the image): one track, tempo 120, 4/4, note_on / note_off pairs of
    events (attack_frame, key, duration_frames, velocity) at MIDI key `key + 21`, velocity round(v / 10 * 127), times
    frame * duration_per_frame converted with mido.second2tick's rounding; messages sorted as the reference sorts its
    (time, type, key, velocity) tuples ('note_off' < 'note_on' at equal times).  Writes a format-1 standard MIDI file
    byte-compatible with what mido saves for the same messages (no running status, end_of_track appended)."""
    def frame_to_tick(frame):
        seconds = frame * duration_per_frame
        return int(round(seconds / (_TEMPO_US * 1e-6 / _TICKS_PER_BEAT)))      # mido.second2tick

    msgs = []
    for attack_frame, key, duration_frame, velocity in events:
        midi_key = int(key) + 21
        vel = int(round((velocity / NUM_VELOCITY_CATEGORIES) * 127))
        msgs.append((frame_to_tick(attack_frame), "note_on", midi_key, vel))
        msgs.append((frame_to_tick(attack_frame + duration_frame), "note_off", midi_key, vel))
    track = bytearray()
    track += _varlen(0) + bytes([0xFF, 0x51, 0x03]) + _TEMPO_US.to_bytes(3, "big")                  # set_tempo
    track += _varlen(0) + bytes([0xFF, 0x58, 0x04, 4, 2, 24, 8])                                    # time_signature 4/4
    now = 0
    for t, kind, key, vel in sorted(msgs):
        track += _varlen(t - now) + bytes([0x90 if kind == "note_on" else 0x80, key & 0x7F, vel & 0x7F])
        now = t
    track += _varlen(0) + bytes([0xFF, 0x2F, 0x00])                                                 # end_of_track
    with open(output_file, "wb") as f:
        f.write(b"MThd" + (6).to_bytes(4, "big") + (1).to_bytes(2, "big") + (1).to_bytes(2, "big") + _TICKS_PER_BEAT.to_bytes(2, "big"))
        f.write(b"MTrk" + len(track).to_bytes(4, "big") + bytes(track))


def read_midi_notes(path: str):
    """Minimal reader of files written by write_midi_file (tests): list of (tick, 'note_on' | 'note_off', key, velocity)."""
    data = open(path, "rb").read()
    assert data[:4] == b"MThd" and data[14:18] == b"MTrk"
    n = int.from_bytes(data[18:22], "big")
    body, i, tick, out = data[22:22 + n], 0, 0, []
    while i < len(body):
        d = 0
        while True:
            b = body[i]
            i += 1
            d = (d << 7) | (b & 0x7F)
            if not b & 0x80:
                break
        tick += d
        st = body[i]
        if st == 0xFF:
            ln = body[i + 2]
            i += 3 + ln
        else:
            out.append((tick, "note_on" if st & 0xF0 == 0x90 else "note_off", body[i + 1], body[i + 2]))
            i += 3
    return out


# ------------------------------------------------------------------------------------------ checkpoints
def save_checkpoint(model, directory: str, step: int):
    """Pytree-path-keyed `.npz` stand-in for the reference's orbax CheckpointManager (train.py:384-394; orbax and
    tensorstore are absent from the image): `<directory>/<step>/params.npz` holds every array leaf under its key path
    (exactly the names and shapes an `ocp.args.StandardRestore` of the reference model yields), `metadata.json` the
    model metadata the reference stores next to a checkpoint (model.py:36-41)."""
    import json
    import os
    from .model import get_model_metadata
    d = os.path.join(directory, str(int(step)))
    os.makedirs(d, exist_ok=True)
    np.savez(os.path.join(d, "params.npz"), **{p: np.asarray(a) for p, a in model.tree_leaves_with_path()})
    with open(os.path.join(d, "metadata.json"), "w") as f:
        json.dump(get_model_metadata(), f)
    return d


def load_newest_checkpoint(checkpoint_path: str):
    """infer.py:172-236 for the `.npz` layout of save_checkpoint: restores the highest step, warns when the stored model
    metadata differs from the current configuration (as the reference does), returns (model, state)."""
    import json
    import os
    from .model import OutputSequenceGenerator, get_model_metadata, model_config
    steps = sorted(int(s) for s in os.listdir(checkpoint_path) if s.isdigit())
    if not steps:
        raise FileNotFoundError("There is no checkpoint to load! Inference will be useless")
    d = os.path.join(checkpoint_path, str(steps[-1]))
    meta_file = os.path.join(d, "metadata.json")
    if os.path.exists(meta_file):
        with open(meta_file) as f:
            stored = json.load(f)
        if stored != json.loads(json.dumps(get_model_metadata())):
            print(f"WARNING: The loaded model has metadata {stored}\nCurrent configuration is {get_model_metadata()}")
    with np.load(os.path.join(d, "params.npz")) as z:
        leaves = {k: z[k] for k in z.files}
    model = OutputSequenceGenerator(model_config, key=1234)
    model.load_leaves(leaves)
    return model, None
