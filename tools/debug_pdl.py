import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from gpu_util import make_model, tap
from oracle import model_np as M, params as P, synth
ACTIVE = dict(gamma_mode="active", decoder_gain=4.0, trained_like=True)
model, tree = make_model(4321, **ACTIVE)
audio = synth.make_windows(2, 4321)
taps = {}
M.forward(P.cast(tree, np.float64), audio[1].astype(np.float64), taps=taps)
a = torch.tensor(audio).cuda()
for trial in range(3):
    got = tap(model, a, "stage4", 2 * 1000 * 64).reshape(2, 1000, 64)[1]
    ref = taps["stage4"]
    err = np.abs(got - ref).max(axis=1)
    bad = np.where(err > 0.05)[0]
    print("trial", trial, "max err", err.max(), "bad rows", len(bad), bad[:40])
got3 = tap(model, a, "stage3", 2 * 2000 * 32).reshape(2, 2000, 32)[1]
print("stage3 err", np.abs(got3 - taps["stage3"]).max())
