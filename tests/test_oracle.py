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


def _piano_roll_probs(rng, frames=400, keys=90, notes=40):
    """Probability track with clear notes (well away from every threshold) plus a low noise floor."""
    p = rng.uniform(0.01, 0.04, size=(frames, keys)).astype(np.float32)
    for _ in range(notes):
        k, a, n = rng.integers(0, keys), rng.integers(0, frames - 30), rng.integers(8, 30)
        t = np.arange(n)
        p[a:a + n, k] = np.maximum(p[a:a + n, k], (0.95 * np.exp(-0.02 * t)).astype(np.float32))
    return p


def test_event_parity_instrument_can_fail():
    """The eventized-parity checker (tests/event_parity.py) used by the GPU tests: (1) its replay of the state machine is the
    oracle's extractor; (2) tracks within the tolerance pass and leave most keys DECIDED; (3) one probability moved across 0.5
    by 2 x TOL fails; (4) a wrong event on a decided key fails; (5) a sub-tolerance nudge across a threshold is excused for
    that key only."""
    import pytest
    from event_parity import check_event_parity, decision_margins
    TOL = 3e-2
    rng = np.random.Generator(np.random.PCG64(11))
    ref = _piano_roll_probs(rng)
    ev, m1, m2 = decision_margins(ref)
    assert ev == E.extract_events(ref) and len(ev) >= 20
    noisy = np.clip(ref + rng.uniform(-TOL / 4, TOL / 4, size=ref.shape).astype(np.float32), 0.0, 1.0)
    n_dec, n_same = check_event_parity(ref, noisy, E.extract_events(noisy), TOL)
    assert n_dec >= 60 and n_same >= n_dec
    a, k = ev[0][0], ev[0][1]                       # the attack frame of a real note: p well above 0.5
    bad = ref.copy()
    bad[a, k] = 0.5 - 2 * TOL                       # (3) across the activation threshold by 2 x TOL
    with pytest.raises(AssertionError, match="tolerance"):
        check_event_parity(ref, bad, E.extract_events(bad), TOL)
    wrong = [e for e in ev if e != ev[0]] + [(ev[0][0] + 1, k, ev[0][2], 7)]      # (4) same probabilities, event shifted a frame
    with pytest.raises(AssertionError, match="event lists differ"):
        check_event_parity(ref, ref, sorted(wrong), TOL)
    near = ref.copy()
    near[:, 3] = 0.02
    near[100:110, 3] = 0.505                        # (5) a plateau 0.005 above the activation threshold ...
    nudged = near.copy()
    nudged[100:110, 3] = 0.495                      # ... pushed just below it: different events, legitimately undecided
    ev_n = E.extract_events(nudged)
    assert [e for e in ev_n if e[1] == 3] != [e for e in E.extract_events(near) if e[1] == 3]
    n_dec2, _ = check_event_parity(near, nudged, ev_n, TOL)
    assert n_dec2 >= 60


def test_autograd_of_twin_matches_fp64_finite_differences():
    """the gradient oracle of the training tests is torch autograd of oracle/model_torch.loss_fn.  Cross-check it
    against central finite differences in fp64 on a shrunken configuration that still has every structure of the path: stem,
    two downsamples, Blocks with active layer scale, one local (16-frame windows over 40 frames, padded to 48) and one global
    transformer layer with RoPE, compressed kv and the GLU feed-forward, decoder, sum-BCE x scale, mean over the batch."""
    conf = {"dims": [4, 8, 16], "depths": [1, 2, 1], "cnn_hidden_expansion": 2.0, "num_transformer_layers": 1,
            "num_transformer_heads": 2, "attention_size": 8, "compressed_attention_q_size": 8, "compressed_attention_kv_size": 4,
            "transformer_dropout_rate": 0.1, "transformer_hidden_expansion": 2.0, "sdd_rate": 0.1}
    tree = P.init_params(3, conf=conf, gamma_mode="active", decoder_gain=2.0, trained_like=True)
    rng = np.random.Generator(np.random.PCG64(5))
    audio = torch.tensor(rng.normal(0, 0.5, size=(2, 2, 800)), dtype=torch.float64)           # 800 / 5 / 2 / 2 = 40 frames
    labels = torch.tensor(np.clip((rng.random((2, 40, 90)) < 0.05).astype(np.float64), 0.005, 0.995))
    cos, sin = T.precompute_frequencies(conf["attention_size"], 64)
    rope = (cos.double(), sin.double())
    tp = T.to_torch(tree, dtype=torch.float64)
    leaves = {}

    def mark(t, prefix=""):
        if isinstance(t, dict):
            return {k: mark(v, f"{prefix}{k}.") for k, v in t.items()}
        if isinstance(t, list):
            return [mark(v, f"{prefix}{i}.") for i, v in enumerate(t)]
        t = t.clone().requires_grad_(True)
        leaves[prefix[:-1]] = t
        return t

    tp = mark(tp)
    loss, logits = T.loss_fn(tp, audio, labels, scale=3.0, rope=rope, conf=conf)
    assert logits.shape == (2, 40, 90)
    loss.backward()
    checked = 0
    worst = 0.0
    with torch.no_grad():
        for path, t in leaves.items():
            if path.endswith("stochastic_depth_dropout.p"):
                assert t.grad is None or float(t.grad.abs().max()) == 0.0     # never reaches the forward (model.py:160,167)
                continue
            flat, g = t.view(-1), t.grad.view(-1)
            for idx in rng.choice(flat.numel(), size=min(2, flat.numel()), replace=False):
                old = float(flat[idx])
                h = 1e-5 * max(1.0, abs(old))
                flat[idx] = old + h
                lp, _ = T.loss_fn(tp, audio, labels, scale=3.0, rope=rope, conf=conf)
                flat[idx] = old - h
                lm, _ = T.loss_fn(tp, audio, labels, scale=3.0, rope=rope, conf=conf)
                flat[idx] = old
                fd = float(lp - lm) / (2 * h)
                err = abs(fd - float(g[idx])) / max(abs(fd), abs(float(g[idx])), 1e-6)
                worst = max(worst, err)
                assert err < 1e-5, (path, int(idx), fd, float(g[idx]))
                checked += 1
    assert checked >= 80
    print(f"finite differences: {checked} entries, worst relative error {worst:.2e}")


def _events_by_bit_masks(probs):
    """numpy restatement of the DEVICE eventizer's algorithm (csrc/event_metrics.cuh): every comparison of the state machine is a
    pure function of the probabilities around a frame, so three masks per key (p > 0.5; p < 0.1; re-attack candidate = p > 0.4,
    not p[f] < p[f + 1], rise of the six-frame means > 0.1) decide everything, and the machine only has to jump from set bit to
    set bit: an idle key to the next `on`, a sounding key to the next `off | re`."""
    p = np.asarray(probs, np.float32)
    F, K = p.shape
    on, off = p > np.float32(0.5), p < np.float32(0.1)
    defer = np.zeros((F, K), bool)
    defer[:-1] = p[:-1] < p[1:]
    rise = np.zeros((F, K), bool)
    for f in range(6, F):
        before = np.zeros(K, np.float32)
        for i in range(f - 6, f):
            before = (before + p[i]).astype(np.float32)          # sequential f32 adds, as the reference forms the mean
        after = np.zeros(K, np.float32)
        for i in range(f, min(f + 6, F)):
            after = (after + p[i]).astype(np.float32)
        rise[f] = (after / np.float32(6.0) - before / np.float32(6.0)).astype(np.float32) > np.float32(0.1)
    re = (p > np.float32(0.4)) & ~defer & rise
    events = []
    for key in range(K):
        on_idx = np.flatnonzero(on[:, key])
        sr_idx = np.flatnonzero(off[:, key] | re[:, key])
        f, started = 0, -1
        while f < F:
            if started < 0:
                i = np.searchsorted(on_idx, f)
                if i == len(on_idx):
                    break
                started = int(on_idx[i])
                f = started + 1
            else:
                i = np.searchsorted(sr_idx, f)
                if i == len(sr_idx):
                    break
                g = int(sr_idx[i])
                if off[g, key]:
                    events.append((started, key, max(g - started, 1), 7))
                    started = -1
                elif g - started > 5:
                    events.append((started, key, max(g - 1 - started, 1), 7))
                    started = g
                f = g + 1
        if started >= 0:
            events.append((started, key, max(F - started, 1), 7))
    return sorted(events)


@pytest.mark.parametrize("frames,kind", [(1, "random"), (7, "random"), (300, "random"), (400, "hover"), (350, "sustained"), (260, "silent")])
def test_bit_mask_walk_equals_the_state_machine(frames, kind):
    """The restructuring the device eventizer rests on (masks + jumps instead of a frame-by-frame state machine) gives the same
    event list as the restatement of common.rs:47-144 -- on the CPU runner, independent of the CUDA code (which
    tests/test_gpu_clip.py holds against the C++ extractor)."""
    from oracle import events as E
    rng = np.random.Generator(np.random.PCG64(frames * 3 + len(kind)))
    K = 24
    if kind == "random":
        p = rng.random((frames, K)).astype(np.float32)
    elif kind == "hover":
        centre = rng.choice(np.float32([0.1, 0.4, 0.5]), size=(1, K))
        p = (centre + rng.normal(0, 0.03, size=(frames, K))).clip(0, 1).astype(np.float32)
        p[::53] = 0.99
    elif kind == "sustained":
        p = (0.3 + 0.6 * rng.random((frames, K))).astype(np.float32)
        p[:, ::3] = np.clip(0.75 + 0.2 * np.sin(np.arange(frames)[:, None] / rng.uniform(3, 40, size=(1, 8))), 0.11, 1).astype(np.float32)
    else:
        p = np.full((frames, K), 0.05, np.float32)
    want = E.extract_events(p)
    assert _events_by_bit_masks(p) == want
    if kind in ("random", "hover", "sustained") and frames >= 300:
        assert len(want) > 20
