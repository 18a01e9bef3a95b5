"""Small profiling target: two plain-stream (no CUDA graph) forwards of 64 windows, so every kernel of the
plan shows up as its own launch under ncu.  Usage: python tools/ncu_target.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import audio_to_midi_b200 as A  # noqa: E402
from oracle import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = A.OutputSequenceGenerator(A.model_config, key=1234)
rope = A.precompute_frequencies(64, 300)
audio = torch.tensor(synth.make_windows_fast(B, 1234), device="cuda:0")
eng = model._engine(0)
eng.L.a2m_set_use_graph(eng.h, 0)
for _ in range(2):
    logits, probs = model.predict(None, audio, rope)
torch.cuda.synchronize()
print("launches per forward:", model.last_launch_count(0), "probs mean", float(probs.mean()))
