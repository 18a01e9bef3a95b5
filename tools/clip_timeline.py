"""Where the time of the long-clip path (config 5) goes on one GPU: python tools/clip_timeline.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_to_midi_b200 as A
from audio_to_midi_b200 import infer as I
from oracle import synth

model = A.OutputSequenceGenerator(A.model_config, key=1234)
base = synth.make_clip(30.0, 1243)
rng = np.random.Generator(np.random.PCG64(1244))
clip = np.concatenate([base * np.float32(g) for g in rng.uniform(0.7, 1.3, size=20)], axis=1).astype(np.float32)
pin = torch.tensor(clip).pin_memory()
rope = A.precompute_frequencies(64, 300)
dev = torch.device("cuda:0")
for it in range(4):
    torch.cuda.synchronize(); t = [time.perf_counter()]
    def mark():
        torch.cuda.synchronize(); t.append(time.perf_counter())
    d = pin.to(dev, non_blocking=True); mark()
    w = I.prepare_windows_device(model, d, 0.5); mark()
    parts = [w[a:b] for a, b in I.balanced_batches(0, int(w.shape[0]), 72)]
    out = model.predict_many(None, parts, rope); mark()
    probs = torch.cat([p for _, p in out]); mark()
    st = I.stitch_probs_device(model, probs, 0.5, 0.02); mark()
    ev = I.extract_events_device(model, st); mark()
    names = ["h2d", "prepare", "forward", "cat", "stitch", "extract"]
    print(it, " ".join(f"{n} {1e3 * (b - a):.2f}" for n, a, b in zip(names, t, t[1:])), "total %.2f ms" % (1e3 * (t[-1] - t[0])), len(ev))
t0 = time.perf_counter(); ev2, _, _ = I.transcribe_clip(model, pin.to(dev, non_blocking=True), overlap=0.5, want_arrays=False); torch.cuda.synchronize()
print("transcribe_clip %.2f ms" % (1e3 * (time.perf_counter() - t0)))
