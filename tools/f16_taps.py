"""Per-tap error of both operand-format variants against the fp32 CPU twin (localises a precision-variant bug to a stage)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import audio_to_midi_b200 as A
from gpu_util import make_model, tap
from oracle import model_torch as T, synth

audio = synth.make_windows(2, 77)
LENS = [16000, 8000, 4000, 2000, 1000, 500, 250]; DIMS = [4, 8, 16, 32, 64, 128, 256]
taps_ref = {}
for precision in ("bf16", "f16"):
    model, tree = make_model(77, precision=precision, gamma_mode="active", decoder_gain=4.0, trained_like=True)
    if not taps_ref:
        with torch.no_grad():
            zref, pref = T.forward(T.to_torch(tree), torch.tensor(audio), taps=taps_ref)
    x = torch.tensor(audio).cuda()
    row = []
    for s in range(7):
        got = tap(model, x, f"stage{s}", 2 * LENS[s] * DIMS[s]).reshape(2, LENS[s], DIMS[s])
        ref = taps_ref[f"stage{s}"].numpy()
        row.append(f"s{s} {np.abs(got - ref).max() / np.abs(ref).max():.1e}")
    for label in ["cnn_out"] + [f"tl{i}_{k}" for i in range(8) for k in ("local", "global")]:
        got = tap(model, x, label, 2 * 256 * 256).reshape(2, 256, 256)[:, :250]
        ref = taps_ref[label].numpy()
        row.append(f"{label} {np.abs(got - ref).max() / np.abs(ref).max():.1e}")
    _, probs = model.predict(None, x, A.precompute_frequencies(64, 300))
    print(precision, "max|dprob| %.3e" % np.abs(probs.cpu().numpy() - pref.numpy()).max(), " ".join(row))
