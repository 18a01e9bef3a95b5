"""-m gpu: the step before the path on the device and the long-audio pipeline of config 5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seconds,overlap,gain", [(23.7, 0.25, 1.0), (12.0, 0.5, 1.0), (7.3, 0.0, 1.0), (9.0, 0.25, 0.001)])
def test_prepare_windows_matches_oracle(seconds, overlap, gain):
    """Normalisation (python.rs:235-264, incl. the quiet-clip branch) + slicing (audio_to_midi_dataset.py:277-294) on the
    device vs the numpy restatement: same window count, values equal up to one f16 rounding tie of the f64 scale."""
    from audio_to_midi_b200 import infer as I
    from gpu_util import make_model
    from oracle import events as E
    rng = np.random.Generator(np.random.PCG64(5))
    n = int(seconds * 16000)
    raw = (rng.normal(0.0, 0.2, (2, n)) * gain).astype(np.float32)
    model, _ = make_model(1)
    got = I.prepare_windows_device(model, raw, overlap).cpu().numpy()
    nl, nr = E.normalize_audio(raw[0], raw[1])
    ref = E.slice_windows(np.stack([nl, nr]), overlap)
    assert got.shape == ref.shape
    # the sum of squares is reduced in a different order (f64): the scale may differ in its last bits, which can flip
    # an f16 rounding tie on a handful of samples at most
    diff = got != ref
    assert diff.mean() < 1e-5
    assert np.abs(got - ref).max() <= np.abs(ref).max() * 2.0 ** -10


def test_transcribe_clip_pipeline():
    """Config 5 in miniature: 32 s synthetic clip -> device normalise + slice -> batched forward -> stitch -> events,
    against the oracle pipeline on the same weights (events compared where no frame is near a threshold)."""
    import audio_to_midi_b200 as A
    from audio_to_midi_b200 import infer as I
    from gpu_util import make_model
    from oracle import events as E
    from oracle import model_torch as T
    from oracle import synth
    model, tree = make_model(99, gamma_mode="active", decoder_gain=4.0, trained_like=True)
    raw = np.asarray(synth.make_clip(32.0, 3), np.float32) * np.float32(0.37)   # un-normalised level
    events, stitched, probs = I.transcribe_clip(model, raw, overlap=0.5, max_batch=4)
    nl, nr = E.normalize_audio(raw[0], raw[1])
    wins = E.slice_windows(np.stack([nl, nr]), 0.5)
    assert probs.shape[0] == wins.shape[0] == 7
    with torch.no_grad():
        _, ref = T.forward(T.to_torch(tree), torch.tensor(wins))
    ref = ref.numpy()
    assert np.abs(probs - ref).max() < 3e-2
    st_ref = E.stitch_probs(ref, 0.5, 0.02)
    assert stitched.shape == st_ref.shape
    assert events == A.modelutil.extract_events(stitched)
    from event_parity import check_event_parity
    nd, ns = check_event_parity(st_ref, stitched, events, 3e-2)
    print(f"32 s clip: decided keys {nd}/90, identical {ns}/90, events {len(events)}")
    assert ns >= nd


def test_validation_loss_and_hit_rate():
    """Config 3 in miniature: per-window BCE sums and event metrics of a batch-partitioned annotated set vs the oracle.
    The event metrics are threshold decisions on the probabilities: for every key whose decisions are all further from a
    threshold than the measured probability difference (tests/event_parity.py) the rasterised prediction column must be
    IDENTICAL to the oracle's, and a window whose 90 keys are all decided must reproduce every metric exactly."""
    import torch.nn.functional as F
    import audio_to_midi_b200 as A
    from audio_to_midi_b200 import infer as I
    from event_parity import decided_keys
    from gpu_util import make_model
    from oracle import events as E
    from oracle import model_torch as T
    from oracle import synth
    model, tree = make_model(99, gamma_mode="active", decoder_gain=4.0, trained_like=True)
    audio, labels = synth.make_windows(6, 11, with_labels=True)
    with torch.no_grad():
        zref, pref = T.forward(T.to_torch(tree), torch.tensor(audio))
        lref = F.binary_cross_entropy_with_logits(zref, torch.tensor(labels), reduction="none").sum(dim=(1, 2)).numpy()
    pref = pref.numpy()
    _, pgpu = model.predict(None, torch.tensor(audio).cuda(), A.precompute_frequencies(64, 300))
    pgpu = pgpu.cpu().numpy()
    assert np.abs(pgpu - pref).max() < 3e-2
    got = {}
    for rank in range(2):                       # two "ranks" in one process: the partition is what is tested
        lo, hi, losses, details = I.compute_testset_loss(model, audio, labels, rank=rank, world_size=2, max_batch=2)
        assert (lo, hi) == ((0, 3), (3, 6))[rank]
        for k in range(hi - lo):
            got[lo + k] = (losses[k], details[k])
    host = I.compute_testset_loss(model, audio, labels, max_batch=4, device_metrics=False)      # host eventizer, one process
    resident = I.compute_testset_loss(model, torch.tensor(audio).cuda(), torch.tensor(labels).cuda(), max_batch=4)   # set already on the device
    assert np.allclose(resident[2], host[2], rtol=1e-5) and len(resident[3]) == 6
    full = 0
    for k in range(6):
        assert abs(got[k][0] - lref[k]) < 5e-3 * abs(lref[k]) + 1.0, (k, got[k][0], lref[k])
        assert abs(host[2][k] - got[k][0]) < 1e-5 * abs(lref[k]) + 1e-3
        ref = E.detailed_event_loss(pref[k], labels[k])
        d = got[k][1]
        assert set(d) == set(ref)
        for q in ref:                            # device metrics == host metrics on the same (GPU) probabilities
            assert abs(d[q] - host[3][k][q]) <= 1e-5 * max(1.0, abs(d[q])), (k, q, d[q], host[3][k][q])
        dec = decided_keys(pref[k], pgpu[k])
        ra = E.to_frame_events(E.extract_events(pgpu[k]), 250)
        rb = E.to_frame_events(E.extract_events(pref[k]), 250)
        assert np.array_equal(ra[:, dec], rb[:, dec]), f"window {k}: a decided key rasterises differently"
        if dec.all():
            full += 1
            for q in ("phantom_notes_diff", "notes_hit"):
                assert d[q] == ref[q], (k, q)
            for q in ("missed_notes_diff", "full_diff", "hit_rate"):
                assert abs(d[q] - ref[q]) <= 1e-5 * max(1.0, abs(ref[q])), (k, q)
        else:                                    # only the undecided columns may differ
            und = ~dec
            slack = float(np.sum(np.abs(ra[:, und] - rb[:, und])))
            assert abs(d["full_diff"] - ref["full_diff"]) <= slack + 1e-3
    print(f"windows with all 90 keys decided: {full}/6")


def test_config5_full_size_properties():
    """BASELINE config 5 at full size -- 600 s clip, overlap 0.5 s -> 134 windows -> 30 175 stitched frames --
    through size-independent properties: the window-partitioned run over 8 "ranks" (17 or 16 windows each) reproduces the single-process probabilities BIT FOR BIT (windows are independent, results do not depend
    on the batch a window travels in), the stitched frame count, and the ordering / range invariants of the event list."""
    import audio_to_midi_b200 as A
    from audio_to_midi_b200 import infer as I
    from gpu_util import make_model
    from oracle import synth
    model, _ = make_model(99, gamma_mode="active", decoder_gain=4.0, trained_like=True)
    raw = np.asarray(synth.make_clip(600.0, 5), np.float32)
    assert raw.shape == (2, 9_600_000)
    events, stitched, probs = I.transcribe_clip(model, raw, overlap=0.5, max_batch=64)
    assert probs.shape == (134, 250, 90)
    assert stitched.shape == (30175, 90)
    assert np.isfinite(probs).all() and probs.min() >= 0.0 and probs.max() <= 1.0
    parts = []
    for rank in range(8):
        lo, hi = I.shard_windows(134, 8, rank)
        assert hi - lo == (17 if rank < 6 else 16)
        _, _, p = I.transcribe_clip(model, raw, overlap=0.5, max_batch=5, rank=rank, world_size=8)
        assert p.shape[0] == hi - lo
        parts.append(p)
    assert np.array_equal(np.concatenate(parts), probs)
    # the event list is the C++ eventizer's: lexicographically sorted (common.rs:142), attacks inside the clip, velocity 7
    assert events == sorted(events)
    assert all(0 <= a < stitched.shape[0] and 0 <= k < 90 and d >= 1 and v == 7 for a, k, d, v in events)


def test_predict_and_stitch_host_windows():
    """infer.py:37-44 (predict_and_stitch) on host windows in several ragged batches (4 + 4 + 2): the pipelined host path
    returns the same probabilities as one device-resident call, bit for bit, and the stitched track of modelutil."""
    import audio_to_midi_b200 as A
    from audio_to_midi_b200 import infer as I
    from gpu_util import make_model
    from oracle import synth
    model, _ = make_model(99, gamma_mode="active", decoder_gain=4.0, trained_like=True)
    wins = synth.make_windows_fast(10, 21)
    probs, stitched, dpf = I.predict_and_stitch(model, None, wins, 5.0, overlap=0.5, max_batch=4)
    _, ref = model.predict(None, torch.tensor(wins).cuda(), A.precompute_frequencies(64, 300))
    assert probs.shape == (10, 250, 90) and abs(dpf - 0.02) < 1e-12
    assert np.array_equal(probs, ref.cpu().numpy())
    assert np.array_equal(stitched, A.modelutil.stitch_probs(probs, 0.5, 0.02))
    # single batch: the plain call path
    probs1, _, _ = I.predict_and_stitch(model, None, wins[:3], 5.0, overlap=0.5, max_batch=4)
    assert np.array_equal(probs1, probs[:3])


@pytest.mark.parametrize("overlap,windows", [(0.5, 9), (0.25, 7), (0.0, 3), (0.49, 5)])
def test_device_stitch_and_extract_bit_exact(overlap, windows):
    """a2m_stitch_probs_dev / a2m_extract_events_dev against the C++ host functions and the oracle on the same probabilities:
    the stitched track is bit-identical (f64 cross-fade, fractional window step of overlap 0.25 -> 12.5 frames, NaN rows of overlap 0),
    the event list is identical, also when the first guess of the event count is too small (the call is repeated)."""
    import audio_to_midi_b200 as A
    from audio_to_midi_b200 import infer as I
    from gpu_util import make_model
    from oracle import events as E
    model, _ = make_model(1)
    rng = np.random.Generator(np.random.PCG64(int(overlap * 100) + windows))
    probs = rng.random((windows, 250, 90)).astype(np.float32) ** 3                     # mostly low, some notes
    for _ in range(60):
        w, k, a, n = rng.integers(0, windows), rng.integers(0, 90), rng.integers(0, 220), rng.integers(8, 30)
        probs[w, a:a + n, k] = np.maximum(probs[w, a:a + n, k], 0.97 * np.exp(-0.02 * np.arange(n)).astype(np.float32))
    host = A.modelutil.stitch_probs(probs, overlap, 0.02)
    dev = I.stitch_probs_device(model, torch.tensor(probs).cuda(), overlap, 0.02)
    assert np.array_equal(dev.cpu().numpy(), host, equal_nan=True)
    assert np.array_equal(host, E.stitch_probs(probs, overlap, 0.02), equal_nan=True)
    clean = np.nan_to_num(host, nan=0.0)                                               # overlap 0 leaves NaN rows (0 / 0, as the reference)
    ev_host = A.modelutil.extract_events(clean)
    ev_dev = I.extract_events_device(model, torch.tensor(clean).cuda())
    assert ev_dev == ev_host and len(ev_host) > 50
    assert ev_host == E.extract_events(clean)
    assert I.extract_events_device(model, torch.tensor(clean).cuda(), cap=2) == ev_host     # too small a buffer -> repeated with the true count


@pytest.mark.parametrize("frames,kind", [(1, "random"), (255, "random"), (256, "random"), (257, "random"), (5000, "random"),
                                         (3000, "sustained"), (3000, "hover"), (2049, "boundaries"), (1500, "silent")])
def test_device_eventizer_bit_mask_walk(frames, kind):
    """a2m_extract_events_dev turns the state machine's comparisons into three bit masks per key (32 frames per word) and lets
    one thread per key jump from set bit to set bit (event_metrics.cuh).  Probability tracks built to stress exactly that against
    the C++ host extractor: tracks shorter than / equal to / one longer than a multiple of the word and tile sizes, keys that never
    fall below 0.1, keys hovering around all four thresholds (re-attack candidates younger than 6 frames must be skipped), notes
    that start or end on word boundaries or sound at the last frame, and an all-silent track."""
    import audio_to_midi_b200 as A
    from audio_to_midi_b200 import infer as I
    from gpu_util import make_model
    model, _ = make_model(1)
    rng = np.random.Generator(np.random.PCG64(frames * 7 + len(kind)))
    if kind == "random":
        p = rng.random((frames, 90)).astype(np.float32)
    elif kind == "sustained":            # never below 0.1; re-attacks (rises > 0.1 over six frames) and long holds spanning segments
        p = (0.3 + 0.6 * rng.random((frames, 90))).astype(np.float32)
        p[:, ::3] = np.clip(0.75 + 0.2 * np.sin(np.arange(frames)[:, None] / rng.uniform(3, 40, size=30)[None, :]), 0.11, 1).astype(np.float32)
    elif kind == "hover":                # small steps around 0.1 / 0.4 / 0.5
        centre = rng.choice(np.float32([0.1, 0.4, 0.5]), size=(1, 90))
        p = (centre + rng.normal(0, 0.03, size=(frames, 90))).clip(0, 1).astype(np.float32)
        p[::97] = 0.99
    elif kind == "boundaries":
        p = np.full((frames, 90), 0.02, dtype=np.float32)
        for k in range(90):
            start = 256 * (1 + k % 7) + (k % 5) - 2            # attacks at seg - 2 .. seg + 2
            stop = 256 * (2 + k % 6) + (k % 3) - 1             # releases at seg - 1 .. seg + 1
            p[start:max(stop, start + 1), k] = 0.9
            p[frames - 1 - (k % 2), k] = 0.95                  # still sounding at the end of the track
    else:
        p = np.full((frames, 90), 0.05, dtype=np.float32)
    want = A.modelutil.extract_events(p)
    got = I.extract_events_device(model, torch.tensor(p).cuda(), cap=16)
    assert got == want
    if kind in ("random", "sustained", "hover", "boundaries") and frames > 256:
        assert len(want) > 50
    if kind == "silent":
        assert want == []
