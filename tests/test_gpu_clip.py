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
    ev_ref = E.extract_events(st_ref)
    fa = A.modelutil.to_frame_events([events], stitched.shape[0])[0]
    fb = A.modelutil.to_frame_events([ev_ref], stitched.shape[0])[0]
    assert np.mean((fa > 0) != (fb > 0)) < 0.05
