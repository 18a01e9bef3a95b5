"""Profiling target for the training plans: one plain-stream (no CUDA graph) forward-with-tape + backward of B windows after
one warm-up, so every kernel shows up as its own launch under ncu.  Usage: python tools/ncu_target_train.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import audio_to_midi_b200 as A  # noqa: E402
from audio_to_midi_b200 import train as T  # noqa: E402
from oracle import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = A.OutputSequenceGenerator(A.model_config, key=1234)
eng = T.TrainEngine(model, 0)
eng.L.a2m_set_use_graph(eng.h, 0)
rope = A.precompute_frequencies(64, 300)
audio = torch.tensor(synth.make_windows_fast(B, 1), device="cuda")
labels = torch.rand(B, 250, 90, device="cuda") * 0.99
eng.set_dropout(0.1, 7)
for _ in range(2):
    eng.zero_grad()
    eng.forward_backward(audio, labels, rope)
torch.cuda.synchronize()
print("launches:", eng.launch_count(), "loss", float(eng.loss.item()))
if len(sys.argv) > 2:   # also one optimizer step (AdamW + clip + re-pack): python tools/ncu_target_train.py 64 opt
    cfg = T.OptimizerConfig()
    for _ in range(2):
        eng.optimizer_step(1e-4, cfg)
    torch.cuda.synchronize()
