"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root:
``python tests/golden/make_golden.py``).

The reference ships no fixtures and cannot be imported here (no JAX), so these
vectors are produced by the fp64 NumPy restatement in oracle/ and pin IT against regressions;
they are not reference outputs.  Inputs are regenerated from seeds (oracle/synth.py,
oracle/params.py); only outputs are stored, as float32, to keep the fixtures small.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from oracle import events as E  # noqa: E402
from oracle import model_np as M  # noqa: E402
from oracle import params as P  # noqa: E402
from oracle import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def forward_case(name, seed, **init_kw):
    p = P.cast(P.init_params(seed, **init_kw), np.float64)
    audio = synth.make_windows(2, seed)
    taps = {}
    rope = M.precompute_frequencies(64, 300, dtype=np.float64)
    logits, probs = M.forward(p, audio[1].astype(np.float64), rope, taps=taps)
    keep = {k: taps[k].astype(np.float32) for k in ("stage4", "stage6", "cnn_out", "tl0_local", "tl0_global", "tl7_global")}
    for k in ("stage0", "stage1", "stage2", "stage3", "stage5"):
        keep[k + "_sub"] = taps[k][::37].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name), logits=logits.astype(np.float32), probs=probs.astype(np.float32),
                        seed=seed, window=1, **keep)
    print(name, logits.shape, float(np.abs(logits).max()))


def local_attention_case():
    rng = np.random.Generator(np.random.PCG64(77))
    p = P.cast(P.init_params(77), np.float64)
    lp = P.layer_slice(p["transformer"]["layers"], 3)["local_attention"]["attention_block"]
    x = rng.normal(size=(250, 256))
    rope = M.precompute_frequencies(64, 300, dtype=np.float64)
    y = M.local_self_attention(x, lp, rope, 4)
    yg = M.self_attention(x, P.layer_slice(p["transformer"]["layers"], 3)["global_attention"]["attention_block"], rope, 4)
    np.savez_compressed(os.path.join(HERE, "attention.npz"), local_out=y.astype(np.float32),
                        global_out=yg.astype(np.float32), seed=77, layer=3)
    print("attention", y.shape)


def events_case():
    rng = np.random.Generator(np.random.PCG64(5))
    # smooth random "probabilities": low-pass noise pushed through a sigmoid, so notes start and stop
    z = rng.normal(size=(3, 250, 90))
    k = np.exp(-np.arange(-12, 13) ** 2 / 30.0)
    z = np.apply_along_axis(lambda v: np.convolve(v, k / k.sum(), mode="same"), 1, z)
    probs = (1.0 / (1.0 + np.exp(-(z * 9.0 - 0.5)))).astype(np.float32)
    out = {"probs": probs}
    for name, ov in (("ov050", 0.5), ("ov025", 0.25), ("ov000", 0.0)):
        st = E.stitch_probs(probs, ov, 0.02)
        out["stitched_" + name] = st
        out["events_" + name] = np.array(E.extract_events(st), dtype=np.int64).reshape(-1, 4)
        out["frames_" + name] = E.to_frame_events([tuple(e) for e in out["events_" + name]], st.shape[0])
    np.savez_compressed(os.path.join(HERE, "events.npz"), **out)
    print("events", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    forward_case("forward_default.npz", 1234)
    forward_case("forward_active.npz", 4321, gamma_mode="active", decoder_gain=4.0, trained_like=True)
    local_attention_case()
    events_case()
