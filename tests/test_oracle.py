"""CPU tests of the oracle itself: golden fixtures, the two restatements against each other,
and the closed-form self-checks of SURVEY.md §4 (the reference has no tests of its own)."""
import os

import numpy as np
import pytest
import torch

from oracle import events as E
from oracle import model_np as M
from oracle import model_torch as T
from oracle import params as P
from oracle import synth


def test_param_count_and_paths():
    p = P.init_params(1)
    assert P.param_count(p) == P.EXPECTED_PARAM_COUNT
    flat = P.flatten(p)
    assert flat["layers.0.layers.0.conv.weight"].shape == (4, 2, 5)
    assert flat["layers.5.layers.21.gamma"].shape == (128,)
    assert flat["layers.6.layers.0.conv.weight"].shape == (256, 128, 2)
    assert flat["transformer.layers.local_attention.attention_block.self_attention.kv_down_proj.weight"].shape == (8, 64, 256)
    assert flat["transformer.layers.global_attention.feed_forward_block.attention_to_intermediate_proj.weight"].shape == (8, 1024, 256)
    assert flat["decoder.decoder_pooling.weight"].shape == (90, 256)
    # field order of Block (model.py:121-126)
    blk = list(p["layers"][2]["layers"][1].keys())
    assert blk == ["depth_conv", "point_conv_1", "point_conv_2", "stochastic_depth_dropout", "norm", "gamma"]


@pytest.mark.parametrize("name,seed,kw", [
    ("forward_default.npz", 1234, {}),
    ("forward_active.npz", 4321, dict(gamma_mode="active", decoder_gain=4.0, trained_like=True)),
])
def test_forward_golden(golden_dir, name, seed, kw):
    g = np.load(os.path.join(golden_dir, name))
    p = P.cast(P.init_params(seed, **kw), np.float64)
    audio = synth.make_windows(2, seed)
    taps = {}
    logits, probs = M.forward(p, audio[1].astype(np.float64), taps=taps)
    np.testing.assert_allclose(logits, g["logits"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(probs, g["probs"], atol=1e-6)
    np.testing.assert_allclose(taps["cnn_out"], g["cnn_out"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(taps["tl7_global"], g["tl7_global"], atol=5e-5, rtol=1e-5)


def test_fp32_twin_matches_fp64():
    kw = dict(gamma_mode="active", decoder_gain=4.0, trained_like=True)
    p = P.init_params(4321, **kw)
    audio = synth.make_windows(2, 4321)
    lg, pr = M.forward_batch(p, audio[1:2])
    with torch.no_grad():
        lt, pt = T.forward(T.to_torch(p), torch.tensor(audio[1:2]))
    assert np.abs(lt.numpy() - lg).max() < 2e-4
    assert np.abs(pt.numpy() - pr).max() < 1e-5


def test_attention_golden_and_index_shift(golden_dir):
    g = np.load(os.path.join(golden_dir, "attention.npz"))
    rng = np.random.Generator(np.random.PCG64(77))
    p = P.cast(P.init_params(77), np.float64)
    layer = P.layer_slice(p["transformer"]["layers"], 3)
    x = rng.normal(size=(250, 256))
    rope = M.precompute_frequencies(64, 300)
    y = M.local_self_attention(x, layer["local_attention"]["attention_block"], rope, 4)
    np.testing.assert_allclose(y, g["local_out"], atol=1e-5)
    yg = M.self_attention(x, layer["global_attention"]["attention_block"], rope, 4)
    np.testing.assert_allclose(yg, g["global_out"], atol=1e-5)
    # index shift (model.py:452-464): rows 0..2 are results of zero-padding tokens.  A padded
    # token has q = k = v = 0, so its attention output is the plain mean of the window's values,
    # and it cannot depend on input rows >= 13 (window 0 covers padded rows 0..15 = tokens -3..12).
    x2 = x.copy()
    x2[13:] += 1.0
    y2 = M.local_self_attention(x2, layer["local_attention"]["attention_block"], rope, 4)
    np.testing.assert_allclose(y2[:3], y[:3], atol=1e-12)
    assert np.abs(y2[8:] - y[8:]).max() > 1e-3
    # row j holds the result for token j-3: perturbing only token 246 (padded row 249) must move
    # output rows 241..249 (windows 29/30 cover padded rows 232..255) and nothing before row 232.
    x3 = x.copy()
    x3[246] += 1.0
    y3 = M.local_self_attention(x3, layer["local_attention"]["attention_block"], rope, 4)
    assert np.abs(y3[:232] - y[:232]).max() < 1e-12
    assert np.abs(y3[249] - y[249]).max() > 1e-4


def test_local_equals_global_when_window_covers_sequence():
    p = P.cast(P.init_params(3), np.float64)
    att = P.layer_slice(p["transformer"]["layers"], 0)["global_attention"]["attention_block"]
    x = np.random.Generator(np.random.PCG64(3)).normal(size=(16, 256))
    rope = M.precompute_frequencies(64, 300)
    a = M.self_attention(x, att, rope, 4)
    b = M.local_self_attention(x, {"self_attention": att}, rope, 4, window=16)
    np.testing.assert_allclose(a, b, atol=1e-12)


def test_block_identity_when_gamma_zero():
    p = P.cast(P.init_params(5), np.float64)
    blk = dict(p["layers"][3]["layers"][1])
    blk["gamma"] = np.zeros_like(blk["gamma"])
    x = np.random.Generator(np.random.PCG64(5)).normal(size=(32, 200))
    assert np.array_equal(M.block(x, blk), x)


def test_events_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "events.npz"))
    for name, ov in (("ov050", 0.5), ("ov025", 0.25), ("ov000", 0.0)):
        st = E.stitch_probs(g["probs"], ov, 0.02)
        assert np.array_equal(st, g["stitched_" + name], equal_nan=True)
    st = g["stitched_ov025"][:300]
    ev = E.extract_events(st)
    ref = [tuple(r) for r in g["events_ov025"].tolist()]
    # events fully decided inside the first 300 frames agree with the full-length run
    assert [e for e in ev if e[0] + e[2] < 280] == [e for e in ref if e[0] + e[2] < 280]


def test_stitch_shapes_and_overlap_zero_nan():
    probs = np.random.Generator(np.random.PCG64(0)).uniform(size=(3, 250, 90)).astype(np.float32)
    assert E.stitch_probs(probs, 0.5, 0.02).shape == (700, 90)          # 25 overlapping frames
    st = E.stitch_probs(probs, 0.25, 0.02)                              # 12.5 overlapping frames
    assert st.shape == (750 - 24, 90)
    np.testing.assert_array_equal(st[:237], probs[0, :237])
    z = E.stitch_probs(probs, 0.0, 0.02)
    assert np.isnan(z[250]).all() and np.isnan(z[500]).all()            # 0/0 blend, as the reference does
    assert np.isfinite(z[251]).all()


def test_extract_roundtrip_well_separated():
    ev = [(10, 5, 30, 7), (60, 5, 20, 7), (15, 40, 100, 7), (200, 89, 50, 7)]
    frames = E.to_frame_events(ev, 250)
    assert E.extract_events(frames) == sorted(ev)


def test_normalize_and_windows():
    rng = np.random.Generator(np.random.PCG64(1))
    l, r = rng.normal(size=1000).astype(np.float32), rng.normal(size=1000).astype(np.float32)
    nl, nr = E.normalize_audio(l, r)
    assert abs(np.sqrt((nl ** 2 + nr ** 2).mean() / 2) - 1.0) < 1e-3
    assert np.array_equal(nl, nl.astype(np.float16).astype(np.float32))
    q = (l * 0.01, r * 0.01)
    ql, _ = E.normalize_audio(*q)
    np.testing.assert_array_equal(ql, q[0].astype(np.float16).astype(np.float32))
    w = E.slice_windows(np.zeros((2, 9_600_000), np.float32), overlap=0.5)
    assert w.shape == (134, 2, 80000)
    assert E.slice_windows(np.zeros((2, 9_600_000), np.float32), overlap=0.25).shape[0] == 127


def test_synth_shapes():
    a, y = synth.make_windows(2, 9, with_labels=True)
    assert a.shape == (2, 2, 80000) and y.shape == (2, 250, 90)
    assert y.min() >= 0.005 and y.max() <= 0.995
    assert np.array_equal(a, a.astype(np.float16).astype(np.float32))


def test_bce_matches_torch():
    rng = np.random.Generator(np.random.PCG64(2))
    z = rng.normal(size=(2, 250, 90)) * 4
    y = rng.uniform(size=(2, 250, 90))
    ref = torch.nn.functional.binary_cross_entropy_with_logits(torch.tensor(z), torch.tensor(y), reduction="none").sum(dim=(1, 2))
    np.testing.assert_allclose(M.bce_with_logits_sum(z, y), ref.numpy(), rtol=1e-12)
