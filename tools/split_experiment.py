"""How should the 134 windows of a 10-minute clip be cut into batches for predict_many's two lanes?  python tools/split_experiment.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_to_midi_b200 as A
from oracle import synth

model = A.OutputSequenceGenerator(A.model_config, key=1234)
rope = A.precompute_frequencies(64, 300)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 134
w = torch.tensor(synth.make_windows_fast(n, 7), device="cuda:0")
def cut(sizes):
    out, i = [], 0
    for s in sizes:
        out.append(w[i:i + s]); i += s
    assert i == n
    return out
def even(k):
    return [n // k + (1 if i < n % k else 0) for i in range(k)]
for sizes in ([64, 64, n - 128], even(3), even(4), even(2), even(6), [n]):
    if min(sizes) <= 0:
        continue
    parts = cut(sizes)
    for _ in range(3):
        model.predict_many(None, parts, rope)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        model.predict_many(None, parts, rope)
    torch.cuda.synchronize()
    print(sizes, "%.3f ms" % (1e2 * (time.perf_counter() - t0)))
