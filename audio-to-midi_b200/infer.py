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
    elif hasattr(samples, "is_cuda") and samples.shape[0] > max_batch:
        # device windows in several batches: two in flight on two streams / workspaces (model.predict_many)
        parts = [samples[i:i + max_batch] for i in range(0, samples.shape[0], max_batch)]
        chunks = [p.cpu().numpy() for _logits, p in model.predict_many(state, parts, rope_freqs)]
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
    _lib.check(eng.h, rc, "a2m_prepare_windows", eng.L)
    return out


def gather_window_blocks(local, n_total: int, world_size: int, rank: int):
    """Rank-ordered concatenation of the per-rank blocks of shard_windows on EVERY rank (contiguous blocks, so the result is
    in window order).  `local` is a torch tensor [n_local, ...] (CUDA -> NCCL, CPU -> gloo).  Blocks differ in size by at most
    one window, so each rank pads its block to the largest one and a single equal-size all_gather moves everything
    (windows x 90 KB: 12 MB for the 10-minute clip of config 5).  Used only AFTER the forward: the path itself has no collective."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return local
    per = [shard_windows(n_total, world_size, r) for r in range(world_size)]
    cap = max(hi - lo for lo, hi in per)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(parts, pad)
    return torch.cat([parts[r][: per[r][1] - per[r][0]] for r in range(world_size)])


def _dist_world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def stitch_probs_device(model, probs, overlap: float, duration_per_frame: float):
    """modelutil.stitch_probs on the device (a2m_stitch_probs_dev, bit-identical to the host function): probs torch CUDA
    [W, F, 90] -> torch CUDA [F', 90].  Falls back to the host function when the overlap is so large that cross-fades chain."""
    import ctypes as C
    import torch
    from . import _lib
    dev = probs.device.index if probs.device.index is not None else torch.cuda.current_device()
    eng = model._engine(dev)
    probs = probs.to(torch.float32).contiguous()
    W, F, K = (int(v) for v in probs.shape)
    stream = C.c_void_p(torch.cuda.current_stream(probs.device).cuda_stream)
    n = int(eng.L.a2m_stitch_probs_dev(eng.h, probs.data_ptr(), W, F, K, float(overlap), float(duration_per_frame), None, 0, stream))
    if n > 0:
        out = torch.empty((n, K), dtype=torch.float32, device=probs.device)
        rc = int(eng.L.a2m_stitch_probs_dev(eng.h, probs.data_ptr(), W, F, K, float(overlap), float(duration_per_frame), out.data_ptr(), n, stream))
        if rc == n:
            return out
    return torch.as_tensor(modelutil.stitch_probs(probs.cpu().numpy(), overlap, duration_per_frame)).to(probs.device)


def extract_events_device(model, stitched, cap: int = 65536):
    """modelutil.extract_events on the device (a2m_extract_events_dev: the comparisons of the hysteresis state machine for all
    (frame, key) pairs in parallel into bit masks, then the machine per key, jumping from set bit to set bit):
    stitched torch CUDA [F, 90] -> the same sorted list of (attack, key, duration, velocity) tuples.  Only the events
    return to the host, one 64-bit word each (attack << 32 | key << 24 | duration), whose ascending order is the order
    common.rs:142 sorts into.  `cap` is the first guess of the event count; the call is repeated with the true count if it
    was too small."""
    import ctypes as C
    import torch
    from . import _lib
    dev = stitched.device.index if stitched.device.index is not None else torch.cuda.current_device()
    eng = model._engine(dev)
    stitched = stitched.to(torch.float32).contiguous()
    F, K = (int(v) for v in stitched.shape)
    if K > 96 or F >= 1 << 24:
        return modelutil.extract_events(stitched.cpu().numpy())
    stream = C.c_void_p(torch.cuda.current_stream(stitched.device).cuda_stream)
    cnt = torch.empty(1, dtype=torch.int32, device=stitched.device)
    while True:
        words = torch.empty(cap, dtype=torch.int64, device=stitched.device)
        rc = eng.L.a2m_extract_events_dev(eng.h, stitched.data_ptr(), F, K, words.data_ptr(), cap, cnt.data_ptr(), stream)
        _lib.check(eng.h, rc, "a2m_extract_events_dev", eng.L)
        n = int(cnt.item())
        if n <= cap:
            break
        cap = n
    if n == 0:
        return []
    w = np.sort(words[:n].cpu().numpy())
    return list(zip((w >> 32).tolist(), ((w >> 24) & 0xFF).tolist(), (w & 0xFFFFFF).tolist(), [7] * n))


def balanced_batches(lo: int, hi: int, max_batch: int) -> list:
    """[(start, stop)] cutting windows lo .. hi into batches of at most max_batch for predict_many's two compute lanes: as few
    batches as max_batch allows, an EVEN number of them when there is more than one (an odd one leaves a lane idle for a whole
    step), and of equal size up to one window -- a forward of 6 windows takes almost as long as one of 64 (every kernel is a
    chain of dependent phases), so a tail batch is the worst cut.  Measured on 134 windows (tools/split_experiment.py):
    64 + 64 + 6 -> 3.64 ms, 67 + 67 -> 2.67 ms."""
    n = hi - lo
    if n <= 0:
        return []
    nb = -(-n // max_batch)
    if nb > 1 and nb % 2:
        nb += 1
    nb = min(nb, n)
    cuts = [lo + (n * i) // nb for i in range(nb + 1)]
    return [(a, b) for a, b in zip(cuts, cuts[1:]) if b > a]


def transcribe_clip(model, audio_samples, overlap: float = 0.25, max_batch: int = 72, rank: int = None, world_size: int = None,
                    gather: bool = True, want_arrays: bool = True):
    """Long-audio transcription (BASELINE config 5; infer.py:339 / audio_to_midi.py:38-53): normalise + slice on the
    device, batched forward of this rank's block of windows, rank-ordered gather of the probabilities, then stitch and
    eventize.  rank / world_size default to the torch.distributed process group (1 process: no collective at all).
    Returns (events, stitched_probs, probs): with several ranks, rank 0 gets the whole clip's events / stitched track and
    every rank the gathered probabilities; with gather=False (or explicit rank / world_size without a process group, as the
    single-process partition tests use) a rank returns (None, None, probabilities of its own block).
    Stitching and event extraction run on the device (stitch_probs_device / extract_events_device, bit-identical to modelutil):
    with want_arrays=False the probabilities never leave the GPU -- only the event list does, which is all audio_to_midi.py:53-56
    needs to write the MIDI file -- and (events, None, None) is returned."""
    import torch
    drank, dworld = _dist_world()
    explicit = rank is not None or world_size is not None
    rank = drank if rank is None else rank
    world_size = dworld if world_size is None else world_size
    windows = prepare_windows_device(model, audio_samples, overlap)
    n_total = int(windows.shape[0])
    lo, hi = shard_windows(n_total, world_size, rank)
    rope_freqs = precompute_frequencies(model_config["attention_size"], 300)
    parts = [windows[a:b] for a, b in balanced_batches(lo, hi, max_batch)]   # max_batch 72: the widest kernels stay one wave of CTAs
    chunks = [p for _lg, p in model.predict_many(None, parts, rope_freqs)]     # consecutive batches overlap on two streams
    local = torch.cat(chunks) if chunks else torch.zeros((0, 250, 90), dtype=torch.float32, device=windows.device)
    if world_size > 1:
        if not gather or (explicit and dworld != world_size):
            return None, None, local.cpu().numpy().astype(np.float32)
        probs_dev = gather_window_blocks(local, n_total, world_size, rank)
        if rank != 0:
            return None, None, (probs_dev.cpu().numpy() if want_arrays else None)
    else:
        probs_dev = local
    stitched_dev = stitch_probs_device(model, probs_dev, overlap, MODEL_AUDIO_LENGTH / probs_dev.shape[1])
    events = extract_events_device(model, stitched_dev)
    if not want_arrays:
        return events, None, None
    return events, stitched_dev.cpu().numpy(), probs_dev.cpu().numpy()


_METRIC_KEYS = ("full_diff", "phantom_notes_diff", "missed_notes_diff", "notes_hit", "hit_rate")


class _HostFeeder:
    """Host arrays -> device, one batch ahead of the forward (validation from numpy arrays, config 3).  A pageable array handed to
    `.to(device)` is copied synchronously and in stream order, i.e. AFTER the previous batch's forward: copy and compute took turns
    (51 ms for 512 windows, 10 k windows/s).  Here the audio and the labels of batch n + 1 are copied into page-locked buffers by
    torch's multi-threaded CPU copy and uploaded on a copy stream while batch n computes; the compute stream only waits for the
    upload's event.  The page-locked ring (two slots) is kept on the model: cudaHostAlloc of 2 x 47 MB costs more than the pass."""

    def __init__(self, model, tdev, max_b: int, shapes):
        import torch
        shapes = [(max_b, *[int(v) for v in sh]) for sh in shapes]
        st = getattr(model, "_eval_ring", None)
        if st is None or st["dev"] != tdev or st["shapes"][0][0] < max_b or [sh[1:] for sh in st["shapes"]] != [sh[1:] for sh in shapes]:
            st = {"dev": tdev, "shapes": shapes, "stream": torch.cuda.Stream(tdev),
                  "pin": [[torch.empty(sh, dtype=torch.float32).pin_memory() for sh in shapes] for _ in range(2)],
                  "buf": [[torch.empty(sh, dtype=torch.float32, device=tdev) for sh in shapes] for _ in range(2)],
                  "uploaded": [None, None],     # event on the copy stream: the uploads out of pin[k] into buf[k] have finished
                  "consumed": [None, None]}     # event on the compute stream: everything that read buf[k] has been passed
            model._eval_ring = st
        self.st, self.tdev, self.rows = st, tdev, {}

    def stage(self, n: int, chunks):
        """Copies the arrays of batch n into the page-locked slot (host threads) and enqueues their uploads on the copy stream."""
        import torch
        st, k = self.st, n % 2
        b = int(chunks[0].shape[0])
        if st["uploaded"][k] is not None:
            st["uploaded"][k].synchronize()                      # the previous uploads have left the page-locked slot
        for pin, chunk in zip(st["pin"][k], chunks):
            pin[:b].copy_(torch.from_numpy(np.ascontiguousarray(chunk, dtype=np.float32)))
        if st["consumed"][k] is not None:
            st["stream"].wait_event(st["consumed"][k])           # the kernels that read buf[k] two batches ago are done
        with torch.cuda.stream(st["stream"]):
            for buf, pin in zip(st["buf"][k], st["pin"][k]):
                buf[:b].copy_(pin[:b], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(st["stream"])
        st["uploaded"][k] = ev
        self.rows[n] = b

    def take(self, n: int):
        """The device tensors of batch n; the current stream waits for their upload."""
        import torch
        st, k = self.st, n % 2
        torch.cuda.current_stream(self.tdev).wait_event(st["uploaded"][k])
        b = self.rows.pop(n)
        return [buf[:b] for buf in st["buf"][k]]

    def consumed(self, n: int):
        """Call after everything that reads batch n's tensors has been enqueued on the current stream."""
        import torch
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.tdev))
        self.st["consumed"][n % 2] = ev


def detailed_event_loss(output_probs: np.ndarray, expected: np.ndarray) -> dict:
    """infer.py:94-158 without the plot: eventize the probabilities, rasterise them back to frames and compare with the
    annotation: full_diff, phantom / missed note mass, notes hit, hit_rate = hit / (hit + phantom + missed).  Host version
    (C++ eventizer), one window at a time as the reference runs it; detailed_event_loss_device is the batched one."""
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


def detailed_event_loss_device(model, probs, expected, want_frames: bool = False):
    """detailed_event_loss for a whole batch in ONE launch on the device (a2m_event_metrics): probs / expected
    torch CUDA tensors [B, F, 90].  Returns the [B, 5] metrics tensor (columns = _METRIC_KEYS) still on the device, and with
    want_frames also the rasterised predictions [B, F, 90]."""
    import ctypes as C
    import torch
    from . import _lib
    dev = probs.device.index if probs.device.index is not None else torch.cuda.current_device()
    eng = model._engine(dev)
    probs = probs.to(torch.float32).contiguous()
    expected = expected.to(probs.device, torch.float32).contiguous()
    B, F, K = probs.shape
    if K != 90 or tuple(expected.shape) != (B, F, K):
        raise ValueError(f"probs / expected must both be (B, F, 90), got {tuple(probs.shape)} / {tuple(expected.shape)}")
    out = torch.empty((B, 5), dtype=torch.float32, device=probs.device)
    frames = torch.empty_like(probs) if want_frames else None
    stream = C.c_void_p(torch.cuda.current_stream(probs.device).cuda_stream)
    rc = eng.L.a2m_event_metrics(eng.h, probs.data_ptr(), expected.data_ptr(), B, F, out.data_ptr(),
                                 frames.data_ptr() if want_frames else None, None, stream)
    _lib.check(eng.h, rc, "a2m_event_metrics", eng.L)
    return (out, frames) if want_frames else out


def metrics_to_dicts(metrics: np.ndarray):
    return [dict(zip(_METRIC_KEYS, (float(v) for v in row))) for row in np.asarray(metrics)]


def compute_testset_loss(model, audio, events, rank: int = None, world_size: int = None, max_batch: int = 64, gather: bool = True,
                         device_metrics: bool = True):
    """Validation pass of config 3 (compute_testset_loss_individual, train.py:86-209; infer.py:94-158): the annotated
    windows are batch-partitioned over ranks (contiguous blocks; the forward has no collective); each rank runs the batched
    forward on its block, the per-window BCE sum (a2m_window_losses) and the event metrics (a2m_event_metrics) on the device --
    nothing but [n, 6] floats ever returns to the host.  With a process group the per-rank results are gathered in rank order.
    audio (N, 2, 80000), events (N, 250, 90): numpy arrays (copied batch by batch), or torch CUDA tensors of a set that is already
    resident on the device.  Returns (lo, hi, losses, [detailed_event_loss dict]): the block bounds
    of this rank, and losses / dicts of ALL windows when gathered (else of the block)."""
    import ctypes as C
    import torch
    from . import _lib
    from .model import _default_device
    drank, dworld = _dist_world()
    explicit = rank is not None or world_size is not None
    rank = drank if rank is None else rank
    world_size = dworld if world_size is None else world_size
    n_total = int(audio.shape[0])
    lo, hi = shard_windows(n_total, world_size, rank)
    dev = _default_device()
    eng = model._engine(dev)
    tdev = torch.device(f"cuda:{dev}")
    rope_freqs = precompute_frequencies(model_config["attention_size"], 300)
    rows = []
    on_device = hasattr(audio, "is_cuda") and audio.is_cuda
    spans = balanced_batches(lo, hi, max_batch)

    def labels_of(i, j):
        if hasattr(events, "is_cuda"):
            return events[i:j].to(tdev, torch.float32).contiguous()
        return torch.as_tensor(np.ascontiguousarray(events[i:j], np.float32)).to(tdev)

    feeder = None
    host_labels = not hasattr(events, "is_cuda")

    def host_batch(i, j):
        return [audio[i:j], events[i:j]] if host_labels else [audio[i:j]]

    if on_device:      # the set is already resident (audio / events torch CUDA tensors): consecutive batches overlap on the two lanes
        outs = model.predict_many(None, [audio[i:j] for i, j in spans], rope_freqs)
    elif spans and dworld == 1 and torch.get_num_threads() >= 4:
        # staging through page-locked memory pays with torch's multi-threaded CPU copy and the host to itself (one process:
        # 10 k -> 27 k windows/s).  Several ranks on one host share its memory bandwidth and run with OMP_NUM_THREADS=1 under
        # torchrun; measured at 2 ranks the staged path was slower than the driver's own pageable copies (8 k against 15-19 k
        # windows/s), so multi-rank processes keep the plain path below
        shapes = [audio.shape[1:], events.shape[1:]] if host_labels else [audio.shape[1:]]
        feeder = _HostFeeder(model, tdev, max(j - i for i, j in spans), shapes)
        feeder.stage(0, host_batch(*spans[0]))
    for n, (i, j) in enumerate(spans):
        if on_device:
            y = labels_of(i, j)
            logits, probs = outs[n]
        elif feeder is None:
            y = labels_of(i, j)
            x = torch.as_tensor(np.ascontiguousarray(audio[i:j], np.float32)).to(tdev)
            logits, probs = model.predict(None, x, rope_freqs)
        else:
            got = feeder.take(n)                                 # the compute stream waits for batch n's upload
            x = got[0]
            y = got[1] if host_labels else labels_of(i, j)
            logits, probs = model.predict(None, x, rope_freqs)
            if n + 1 < len(spans):                               # batch n + 1 is staged and uploaded while batch n computes
                feeder.stage(n + 1, host_batch(*spans[n + 1]))
        out = torch.empty(j - i, dtype=torch.float32, device=tdev)
        stream = C.c_void_p(torch.cuda.current_stream(tdev).cuda_stream)
        _lib.check(eng.h, eng.L.a2m_window_losses(eng.h, logits.data_ptr(), y.data_ptr(), j - i, out.data_ptr(), stream), "a2m_window_losses", eng.L)
        if device_metrics:
            m = detailed_event_loss_device(model, probs, y)
        else:
            pr = probs.cpu().numpy()
            ev_host = events[i:j].cpu().numpy() if hasattr(events, "is_cuda") else events[i:j]
            m = torch.tensor([[detailed_event_loss(pr[k], ev_host[k])[q] for q in _METRIC_KEYS] for k in range(j - i)],
                             dtype=torch.float32, device=tdev)
        rows.append(torch.cat([out[:, None], m], dim=1))
        if feeder is not None:
            feeder.consumed(n)
    local = torch.cat(rows) if rows else torch.zeros((0, 6), dtype=torch.float32, device=tdev)
    if world_size > 1 and gather and not (explicit and dworld != world_size):
        local = gather_window_blocks(local, n_total, world_size, rank)
    table = local.cpu().numpy()
    return lo, hi, table[:, 0].copy(), metrics_to_dicts(table[:, 1:])


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
def save_checkpoint(model, directory: str, step: int, ensemble_axis: bool = True):
    """Pytree-path-keyed `.npz` stand-in for the reference's orbax CheckpointManager (train.py:384-394; orbax and
    tensorstore are absent from the image): `<directory>/<step>/params.npz` holds every array leaf under its key path,
    `metadata.json` the model metadata the reference stores next to a checkpoint (model.py:36-41).

    ensemble_axis (default): every array carries the reference's leading ENSEMBLE axis of size 1 -- the reference trains a
    `filter_vmap`-ed ensemble (train.py:788-795) and saves `eqx.filter(model_ensemble, is_inexact_array)`, so a checkpointed
    leaf is (1, ...) and a transformer leaf (1, 8, ...).  These are exactly the names and shapes an
    `ocp.args.StandardRestore` of the reference model yields, which is what tools/convert_orbax.py moves in and out of orbax."""
    import json
    import os
    from .model import get_model_metadata
    d = os.path.join(directory, str(int(step)))
    os.makedirs(d, exist_ok=True)
    leaves = {p: np.asarray(a) for p, a in model.tree_leaves_with_path()}
    if ensemble_axis:
        leaves = {p: a[None, ...] for p, a in leaves.items()}
    np.savez(os.path.join(d, "params.npz"), **leaves)
    with open(os.path.join(d, "metadata.json"), "w") as f:
        json.dump({**get_model_metadata(), "ensemble_axis": bool(ensemble_axis)}, f)
    return d


def select_ensemble_member(leaves: dict, reference_shapes: dict, ensemble_select: int = 0) -> dict:
    """infer.py:213-221 (ensemble_selector): checkpointed arrays carry a leading ensemble axis; pick member
    `ensemble_select`.  Arrays that already have the model's shape (a checkpoint written without the axis) pass through."""
    out = {}
    for p, a in leaves.items():
        a = np.asarray(a)
        want = tuple(reference_shapes[p]) if p in reference_shapes else None
        if want is not None and a.shape == want:
            out[p] = a
        elif want is not None and a.ndim == len(want) + 1 and a.shape[1:] == want:
            if not 0 <= ensemble_select < a.shape[0]:
                raise IndexError(f"leaf {p}: ensemble member {ensemble_select} of {a.shape[0]}")
            out[p] = a[ensemble_select]
        else:
            out[p] = a          # load_leaves reports the shape mismatch with the leaf's name
    return out


def load_newest_checkpoint(checkpoint_path: str, ensemble_size: int = 1, ensemble_select: int = 0):
    """infer.py:172-236 for the `.npz` layout of save_checkpoint: restores the highest step, warns when the stored model
    metadata differs from the current configuration (as the reference does), selects ensemble member `ensemble_select`
    (infer.py:213-221), returns (model, state)."""
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
        stored.pop("ensemble_axis", None)
        if stored != json.loads(json.dumps(get_model_metadata())):
            print(f"WARNING: The loaded model has metadata {stored}\nCurrent configuration is {get_model_metadata()}")
    with np.load(os.path.join(d, "params.npz")) as z:
        leaves = {k: z[k] for k in z.files}
    model = OutputSequenceGenerator(model_config, key=1234)
    shapes = {p: np.shape(a) for p, a in model.tree_leaves_with_path()}
    model.load_leaves(select_ensemble_member(leaves, shapes, ensemble_select))
    return model, None
