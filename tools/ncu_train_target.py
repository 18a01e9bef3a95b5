"""Profiling target: two training steps (forward with tape + backward + AdamW) of B windows on plain streams."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_to_midi_b200 as A
from audio_to_midi_b200 import train as T
from oracle import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = A.OutputSequenceGenerator(A.model_config, key=1234)
eng = T.TrainEngine(model, 0)
rope = A.precompute_frequencies(64, 300)
audio = torch.tensor(synth.make_windows_fast(B, 1), device="cuda")
labels = torch.rand(B, 250, 90, device="cuda") * 0.99
cfg = T.OptimizerConfig()
for _ in range(2):
    loss, ok, _ = eng.training_step(audio, labels, rope, cfg, 1e-4)
torch.cuda.synchronize()
print("launches per step:", eng.launch_count(), "loss", float(loss.item()))
