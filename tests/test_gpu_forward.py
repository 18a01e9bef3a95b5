"""Forward parity of the CUDA path against the fp64 oracle, through the C ABI.

Tolerances (stated per SURVEY.md §7 "fp16 vs bf16"): stage 0 is an fp32 CUDA-core kernel with the hardware tanh in its GELU (1e-3); from stage 1 on
every pointwise contraction has bf16 operands with fp32 accumulation (tcgen05), the residual stream is fp32.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ACTIVE = dict(gamma_mode="active", decoder_gain=4.0, trained_like=True)
LENS = [16000, 8000, 4000, 2000, 1000, 500, 250]
DIMS = [4, 8, 16, 32, 64, 128, 256]


@pytest.fixture(scope="module")
def setup():
    from gpu_util import make_model
    from oracle import model_np as M
    from oracle import params as P
    from oracle import synth
    model, tree = make_model(4321, **ACTIVE)
    audio = synth.make_windows(2, 4321)
    taps = {}
    p64 = P.cast(tree, np.float64)
    logits, probs = M.forward(p64, audio[1].astype(np.float64), taps=taps)
    return model, tree, audio, taps, logits, probs


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-6)


@pytest.mark.parametrize("stage", range(7))
def test_cnn_stage_taps(setup, stage):
    from gpu_util import tap
    model, _, audio, taps, _, _ = setup
    a = torch.tensor(audio).cuda()
    got = tap(model, a, f"stage{stage}", 2 * LENS[stage] * DIMS[stage]).reshape(2, LENS[stage], DIMS[stage])[1]
    ref = taps[f"stage{stage}"]
    tol = 1e-3 if stage == 0 else 3e-2
    assert _rel(got, ref) < tol, f"stage {stage}: rel err {_rel(got, ref)}"


@pytest.mark.parametrize("label", ["cnn_out", "tl0_local", "tl0_global", "tl3_local", "tl7_global"])
def test_transformer_taps(setup, label):
    from gpu_util import tap
    model, _, audio, taps, _, _ = setup
    a = torch.tensor(audio).cuda()
    got = tap(model, a, label, 2 * 256 * 256).reshape(2, 256, 256)[1, :250]
    ref = taps[label]
    assert _rel(got, ref) < 4e-2, f"{label}: rel err {_rel(got, ref)}"


def test_forward_probs_and_events(setup):
    import audio_to_midi_b200 as A
    model, _, audio, _, logits, probs = setup
    rope = A.precompute_frequencies(64, 300)
    lg, pr = model.predict(None, torch.tensor(audio).cuda(), rope)
    lg, pr = lg.cpu().numpy()[1], pr.cpu().numpy()[1]
    tol = 3e-2                                           # bf16 tensor path, probabilities
    assert np.abs(pr - probs).max() < tol, np.abs(pr - probs).max()
    assert np.abs(lg - logits).max() < 0.15, np.abs(lg - logits).max()
    assert np.allclose(pr, 1 / (1 + np.exp(-lg.astype(np.float64))), atol=2e-6)
    # host path (numpy in/out) gives the same numbers as the device path
    lg_h, pr_h = model.predict(None, audio, rope)
    assert np.array_equal(lg_h[1], lg) and np.array_equal(pr_h[1], pr)
    # single-window call shape of the reference (2, N) -> (250, 90)
    lg1, _ = model.predict(None, audio[1], rope)
    assert lg1.shape == (250, 90) and np.abs(lg1 - lg).max() < 1e-5


def test_default_init_forward(golden_dir):
    """gamma = 1e-6 (reference default init): checked against the committed golden vector."""
    import os
    import audio_to_midi_b200 as A
    from gpu_util import make_model
    from oracle import synth
    g = np.load(os.path.join(golden_dir, "forward_default.npz"))
    model, _ = make_model(1234)
    audio = synth.make_windows(2, 1234)
    _, pr = model.predict(None, audio, A.precompute_frequencies(64, 300))
    assert np.abs(pr[1] - g["probs"]).max() < 2e-2


def test_pipelined_host_path_matches_sync(setup):
    """predict_pipelined (submit/collect on two slots, pinned buffers) returns the same numbers as predict."""
    import audio_to_midi_b200 as A
    model, _, audio, _, _, _ = setup
    rope = A.precompute_frequencies(64, 300)
    rng = np.random.Generator(np.random.PCG64(5))
    batches = []
    for i in range(5):
        b = A.pinned_empty((2 + (i % 2), 2, 80000))          # ragged batch sizes
        b[...] = audio[:1] * rng.uniform(0.5, 1.5, size=(b.shape[0], 1, 1)).astype(np.float32)
        batches.append(b)
    outs = list(model.predict_pipelined(iter(batches), rope))
    assert len(outs) == 5
    for b, (lg, pr) in zip(batches, outs):
        lg_s, pr_s = model.predict(None, np.array(b), rope)
        assert lg.shape == (b.shape[0], 250, 90)
        assert np.array_equal(lg, lg_s) and np.array_equal(pr, pr_s)
    assert list(model.predict_pipelined(iter([]), rope)) == []


def test_odd_batch_matches_torch_oracle():
    """5 windows (an odd number of 128-row tiles per CNN stage, 10 transformer tiles): every window against the
    PyTorch-CPU fp32 restatement, so tile boundaries inside and between windows are exercised in the fused kernels."""
    import audio_to_midi_b200 as A
    from gpu_util import make_model
    from oracle import model_torch as MT
    from oracle import synth
    model, tree = make_model(97, **ACTIVE)
    audio = synth.make_windows(5, 97)
    with torch.no_grad():
        ref_logits, ref_probs = MT.forward(MT.to_torch(tree), torch.tensor(audio))
    _, pr = model.predict(None, torch.tensor(audio).cuda(), A.precompute_frequencies(64, 300))
    err = np.abs(pr.cpu().numpy() - ref_probs.numpy()).reshape(5, -1).max(axis=1)
    assert (err < 3e-2).all(), err


def test_fused_and_unfused_paths_agree(tmp_path):
    """The run-time switches of INTEGRATION.md section 5 select launch-by-launch equivalents of the fused kernels
    (read once at a2m_create, hence one subprocess per setting): every setting must reproduce the default path's
    probabilities to bf16 accuracy.  Guards both the switches and the fused kernels against each other."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = (
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "import audio_to_midi_b200 as A\n"
        "from gpu_util import make_model\n"
        "from oracle import synth\n"
        "model, _ = make_model(4321, gamma_mode='active', decoder_gain=4.0, trained_like=True)\n"
        "audio = synth.make_windows(3, 4321)\n"
        "_, pr = model.predict(None, torch.tensor(audio).cuda(), A.precompute_frequencies(64, 300))\n"
        "np.save(sys.argv[1], pr.cpu().numpy())\n")
    outs = {}
    settings = {"default": {}, "unfused": {"A2M_FUSE_QKV": "0", "A2M_FUSE_FFN": "0", "A2M_FUSE_B256": "0", "A2M_FUSE_SMALL": "0"},
                "no_graph_no_pdl": {"A2M_GRAPH": "0", "A2M_PDL": "0"}}
    for name, env in settings.items():
        out = tmp_path / f"{name}.npy"
        res = subprocess.run([sys.executable, "-c", script, str(out)], env={**os.environ, **env}, capture_output=True, text=True,
                             timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        outs[name] = np.load(out)
    assert np.array_equal(outs["default"], outs["no_graph_no_pdl"])       # same kernels, different launch mechanism
    d = np.abs(outs["unfused"] - outs["default"]).max()       # LN / GEMM / GLU launch by launch instead of the fused kernels
    assert 0 < d < 2e-2, d


def test_batch_invariance_bitwise():
    """Windows are independent (infer.py:40 is a pure vmap): the probabilities of a window must not depend on the batch it
    travels in -- bit for bit, across batch sizes that change every kernel's grid (97 = one call vs 64 + 33 vs 97 x 1's
    first and last).  This is the property the window-partitioned multi-GPU paths (configs 3 and 5) rest on."""
    import audio_to_midi_b200 as A
    from gpu_util import make_model
    from oracle import synth
    model, _ = make_model(4321, gamma_mode="active", decoder_gain=4.0, trained_like=True)
    rope = A.precompute_frequencies(64, 300)
    audio = torch.tensor(synth.make_windows_fast(97, 77)).cuda()
    lg, pr = model.predict(None, audio, rope)
    lg_a, pr_a = model.predict(None, audio[:64], rope)
    lg_b, pr_b = model.predict(None, audio[64:], rope)
    assert torch.equal(pr, torch.cat([pr_a, pr_b])) and torch.equal(lg, torch.cat([lg_a, lg_b]))
    for k in (0, 96):
        _, p1 = model.predict(None, audio[k], rope)
        assert torch.equal(p1, pr[k])
    assert float(pr.std()) > 1e-3


def test_host_path_f16_in_probs_only_f16_out(setup):
    """a2m_submit_host_ex: audio shipped as IEEE binary16 (lossless: load_full_audio rounds to f16, python.rs:235-264), logits not
    returned (infer.py:41 keeps only probs), probabilities returned as f32 or f16.  Same kernels, so the f32 probabilities are
    bit-identical to the all-f32 call and the f16 ones are their correctly rounded values."""
    import audio_to_midi_b200 as A
    model, _tree, audio = setup[0], setup[1], setup[2]
    rope = A.precompute_frequencies(64, 300)
    a32 = np.ascontiguousarray(audio[:4])
    a16 = a32.astype(np.float16)
    assert np.array_equal(a16.astype(np.float32), a32)             # synthetic windows are f16-exact by construction
    (lg, pr), = list(model.predict_pipelined([a32], rope, copy=True))
    (lg2, pr2), = list(model.predict_pipelined([a16], rope, copy=True, want_logits=False))
    assert lg2 is None and np.array_equal(pr2, pr)
    (lg3, pr3), = list(model.predict_pipelined([a16], rope, copy=True, want_logits=False, probs_dtype=np.float16))
    assert pr3.dtype == np.float16 and np.array_equal(pr3, pr.astype(np.float16))
    (lg4, pr4), = list(model.predict_pipelined([a32], rope, copy=True, want_logits=True, probs_dtype=np.float16))
    assert np.array_equal(lg4, lg) and np.array_equal(pr4, pr3)


def test_f16_operands_tighten_parity():
    """a17 / change_fp_precision (infer.py:27-32, 234): the same kernels built with IEEE binary16 tensor-core operands
    (libaudio2midi_b200_f16.so) against the fp32 CPU twin, next to the bf16 build, on 8 windows of the "active" model (every
    Block matters, probabilities spread over (0, 1)).  binary16 carries 3 more significand bits than bf16, so the probability
    error must drop several-fold; both variants launch the same plan and both residual taps agree with the twin."""
    import audio_to_midi_b200 as A
    from gpu_util import make_model, tap
    from oracle import model_torch as T
    from oracle import synth
    audio = synth.make_windows(8, 77)
    rope = A.precompute_frequencies(64, 300)
    errs, launches = {}, {}
    taps_ref = {}
    for precision in ("bf16", "f16"):
        model, tree = make_model(77, precision=precision, **ACTIVE)
        if not taps_ref:
            with torch.no_grad():
                zref, pref = T.forward(T.to_torch(tree), torch.tensor(audio), taps=taps_ref)
            pref = pref.numpy()
        _, probs = model.predict(None, torch.tensor(audio).cuda(), rope)
        probs = probs.cpu().numpy()
        errs[precision] = float(np.abs(probs - pref).max())
        launches[precision] = model.last_launch_count(0)
        assert model._engine(0).precision == precision
        for label, shape in (("stage5", (8 * 500, 128)), ("tl7_global", (8 * 256, 256))):
            got = tap(model, torch.tensor(audio).cuda(), label, shape[0] * shape[1]).reshape(shape)
            ref = taps_ref[label].numpy()
            if label.startswith("tl"):
                got = got.reshape(8, 256, 256)[:, :250]
            else:
                got = got.reshape(8, 500, 128)
            rel = _rel(got, ref)
            assert rel < (4e-2 if precision == "bf16" else 4e-3), (precision, label, rel)     # measured 1.0e-2 / 1.0e-3
    print(f"max |dprob| vs the fp32 twin: bf16 operands {errs['bf16']:.3e}, f16 operands {errs['f16']:.3e}")
    assert launches["bf16"] == launches["f16"]
    assert errs["bf16"] < 3e-2
    assert errs["f16"] < 5e-3 and errs["f16"] < 0.25 * errs["bf16"]                        # measured 2.0e-2 / 2.4e-3


def test_predict_many_overlaps_batches_bitwise():
    """model.predict_many: consecutive independent batches alternate over two streams / workspaces / launch plans.  Results are
    bit-identical to one predict call per batch (windows are independent and each lane runs the same kernels), for ragged batch
    sizes, repeated calls, and with work enqueued on the caller's stream before and after (stream-ordered, no host sync)."""
    import audio_to_midi_b200 as A
    from gpu_util import make_model
    from oracle import synth
    model, _ = make_model(4321, **ACTIVE)
    rope = A.precompute_frequencies(64, 300)
    wins = torch.tensor(synth.make_windows_fast(23, 5)).cuda()
    parts = [wins[0:8], wins[8:16], wins[16:21], wins[21:23], wins[0:8]]
    ref = [model.predict(None, p, rope) for p in parts]
    torch.cuda.synchronize()
    for _ in range(3):
        scaled = [p * 1.0 for p in parts]                      # produced on the current stream right before the call
        got = model.predict_many(None, scaled, rope)
        total = sum(float(pr.sum()) for _, pr in got)          # consumed on the current stream right after it
        assert np.isfinite(total)
        assert len(got) == len(ref)
        for (lg, pr), (rl, rp) in zip(got, ref):
            assert torch.equal(lg, rl) and torch.equal(pr, rp)
    one = model.predict_many(None, [wins[:3]], rope)
    assert torch.equal(one[0][1], model.predict(None, wins[:3], rope)[1])


@pytest.mark.gpu
def test_empty_batches_on_the_device():
    """A zero-length batch gives zero-length outputs on the device path too, alone and inside predict_many (a rank whose block of a
    short clip is empty), without disturbing its neighbours."""
    import audio_to_midi_b200 as A
    from gpu_util import make_model
    from oracle import synth
    model, _ = make_model(4321, **ACTIVE)
    rope = A.precompute_frequencies(64, 300)
    wins = torch.tensor(synth.make_windows_fast(5, 9)).cuda()
    lg, pr = model.predict(None, wins[:0], rope)
    assert tuple(lg.shape) == tuple(pr.shape) == (0, 250, 90) and pr.is_cuda
    got = model.predict_many(None, [wins[:2], wins[:0], wins[2:5]], rope)
    assert tuple(got[1][1].shape) == (0, 250, 90)
    assert torch.equal(got[0][1], model.predict(None, wins[:2], rope)[1])
    assert torch.equal(got[2][1], model.predict(None, wins[2:5], rope)[1])


@pytest.mark.gpu
def test_large_batch_matches_small_batches_bitwise():
    """260 windows in one call: every kernel runs several waves of CTAs (stage 5: 1016 tiles on 296 slots) and the workspace is
    four times the bench's.  Its probabilities equal, bit for bit, those of the same windows sent in batches of 64 / 64 / 64 / 68."""
    import audio_to_midi_b200 as A
    from gpu_util import make_model
    from oracle import synth
    model, _ = make_model(4321, **ACTIVE)
    rope = A.precompute_frequencies(64, 300)
    base = torch.tensor(synth.make_windows_fast(65, 21)).cuda()
    audio = torch.cat([base * g for g in (1.0, 0.83, 1.21, 0.67)])           # 260 distinct windows
    _, pr = model.predict(None, audio, rope)
    parts = torch.cat([model.predict(None, audio[a:b], rope)[1] for a, b in ((0, 64), (64, 128), (128, 192), (192, 260))])
    assert torch.equal(pr, parts)
    assert bool(torch.isfinite(pr).all()) and float(pr.std()) > 1e-3
