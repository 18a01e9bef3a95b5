"""Eventized parity (BASELINE config 2, at its full size of 64 windows): the thresholded MIDI event lists computed from the
CUDA path's probabilities vs the oracle's, with the decision-margin instrument of tests/event_parity.py -- a key's event list
MUST be identical whenever every comparison the extractor evaluated on that key is further from flipping than the measured
probability difference.  Nothing is excused globally; the number of keys the assertion had force on is pinned."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# probability tolerance against the fp32 CPU twin, per tensor-core operand format (tests/test_gpu_forward.py measures both)
TOLS = {"bf16": 3e-2, "f16": 4e-3}    # measured maxima over 64 windows: 1.8e-2 / 2.1e-3
N_WINDOWS = 64
# Keys (of 64 windows x 90) whose event list the margin theorem FORCES to be identical, measured on B200 in round 2 with the
# probability differences of each variant; the bounds sit a little below the measurement so that the assertion cannot go vacuous.
MIN_DECIDED = {"bf16": 600, "f16": 1150}            # measured 657 / 1244 of 5760 (identical: 4750 / 5629)
MIN_DECIDED_STITCHED = {"bf16": 0, "f16": 0}


@pytest.fixture(scope="module", params=["bf16", "f16"])
def run64(request):
    import audio_to_midi_b200 as A
    from gpu_util import make_model
    from oracle import model_torch as T
    from oracle import synth
    model, tree = make_model(99, precision=request.param, gamma_mode="active", decoder_gain=4.0, trained_like=True)
    audio = synth.make_windows(N_WINDOWS, 99)
    rope = A.precompute_frequencies(64, 300)
    _, probs = model.predict(None, torch.tensor(audio).cuda(), rope)
    probs = probs.cpu().numpy()
    with torch.no_grad():
        tp = T.to_torch(tree)
        ref = np.concatenate([T.forward(tp, torch.tensor(audio[i:i + 8]))[1].numpy() for i in range(0, N_WINDOWS, 8)])
    return model, probs, ref


def test_probabilities_of_64_windows_within_tolerance(run64):
    model, probs, ref = run64
    d = np.abs(probs.astype(np.float64) - ref)
    print(f"[{model.precision}] 64 windows: max |dprob| {d.max():.3e}, mean {d.mean():.3e}, p99.9 {np.quantile(d, 0.999):.3e}")
    assert d.max() < TOLS[model.precision]


def test_event_lists_match_where_unambiguous(run64):
    """Per window (250 frames x 90 keys each, 64 x 90 = 5760 key tracks) and for the stitched track of all 64 windows."""
    import audio_to_midi_b200 as A
    from event_parity import check_event_parity
    from oracle import events as E
    model, probs, ref = run64
    TOL = TOLS[model.precision]
    decided = same = total_events = 0
    for w in range(N_WINDOWS):
        ev = A.modelutil.extract_events(np.ascontiguousarray(probs[w]))
        assert ev == E.extract_events(probs[w])                      # same probabilities: C++ extractor == oracle extractor
        nd, ns = check_event_parity(ref[w], probs[w], ev, TOL)
        decided += nd
        same += ns
        total_events += len(ev)
    st = A.modelutil.stitch_probs(probs, 0.5, 0.02)
    assert np.array_equal(st, E.stitch_probs(probs, 0.5, 0.02), equal_nan=True)
    st_ref = E.stitch_probs(ref, 0.5, 0.02)
    ev_st = A.modelutil.extract_events(st)
    nd_st, ns_st = check_event_parity(st_ref, st, ev_st, TOL)
    print(f"[{model.precision}] per-window key tracks: decided {decided}/5760, identical {same}/5760, events {total_events}; "
          f"stitched ({st.shape[0]} frames): decided {nd_st}/90, identical {ns_st}/90, events {len(ev_st)}")
    # measured on B200 (round 2, bf16 operands): see DESIGN.md section 5; the bounds keep the assertion from going vacuous
    assert total_events > 500
    assert decided >= MIN_DECIDED[model.precision] and same >= decided
    assert nd_st >= MIN_DECIDED_STITCHED[model.precision]


def test_event_metrics_on_device_match_host(run64):
    """a2m_event_metrics (one launch, one thread per (window, key)) vs the oracle's detailed_event_loss on the SAME
    probabilities: counts exact, float sums to 1e-5; rasterised predictions bit-identical to modelutil.to_frame_events."""
    import audio_to_midi_b200 as A
    from audio_to_midi_b200 import infer as I
    from oracle import events as E
    from oracle import synth
    model, probs, _ = run64
    _, labels = synth.make_windows(N_WINDOWS, 99, with_labels=True)
    m, frames = I.detailed_event_loss_device(model, torch.tensor(probs).cuda(), torch.tensor(labels).cuda(), want_frames=True)
    torch.cuda.synchronize()
    m, frames = m.cpu().numpy(), frames.cpu().numpy()
    for w in range(0, N_WINDOWS, 3):
        refd = E.detailed_event_loss(probs[w], labels[w])
        got = I.metrics_to_dicts(m[w:w + 1])[0]
        host = I.detailed_event_loss(probs[w], labels[w])
        for d in (got, host):
            assert d["phantom_notes_diff"] == refd["phantom_notes_diff"] and d["notes_hit"] == refd["notes_hit"]
            assert abs(d["missed_notes_diff"] - refd["missed_notes_diff"]) <= 1e-5 * max(1.0, refd["missed_notes_diff"])
            assert abs(d["full_diff"] - refd["full_diff"]) <= 1e-5 * max(1.0, refd["full_diff"])
            assert abs(d["hit_rate"] - refd["hit_rate"]) <= 1e-6
        raster = A.modelutil.to_frame_events([A.modelutil.extract_events(np.ascontiguousarray(probs[w]))], 250)[0]
        assert np.array_equal(frames[w], raster)
        assert np.array_equal(raster, E.to_frame_events(E.extract_events(probs[w]), 250))
