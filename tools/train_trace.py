"""Trace of a few training steps (loss, finite flag, update norm) with the bench's settings, to check that the loss
stays finite and decreases.  Usage: python tools/train_trace.py [batch] [steps] [dropout]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import audio_to_midi_b200 as A  # noqa: E402
from audio_to_midi_b200 import train as T  # noqa: E402
from oracle import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 12
DO = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
dev = torch.device("cuda:0")
model = A.OutputSequenceGenerator(A.model_config, key=1234)
eng = T.TrainEngine(model, 0)
cfg = T.OptimizerConfig()
eng.set_lr_multipliers(T.layer_lr_multipliers(eng.paths, cfg.layer_lr_decay))
sched = T.create_learning_rate_schedule(cfg.base_learning_rate, cfg.warmup_steps, cfg.num_steps)
rope = A.precompute_frequencies(64, 300)
x = torch.tensor(synth.make_windows_fast(B, 1234), device=dev)
rng = np.random.Generator(np.random.PCG64(1234))
y = torch.tensor(np.clip((rng.random((B, 250, 90)) < 0.02).astype(np.float32), 0.005, 0.995), device=dev)
for i in range(K):
    loss, valid, _ = eng.training_step(x, y, rope, cfg, sched(i + 1), dropout_rate=DO, key=1234)
    torch.cuda.synchronize()
    g = eng.grads
    p = eng.params_flat()
    print(f"step {i}: loss {float(loss.item()):.4f} valid {bool(valid.item())} stats {eng.stats.tolist()} "
          f"|g| {float(g.norm()):.4e} g_nonfinite {int((~torch.isfinite(g)).sum())} p_nonfinite {int((~torch.isfinite(p)).sum())}")
    if int((~torch.isfinite(g)).sum()):
        bad = (~torch.isfinite(g)).nonzero().flatten().cpu().numpy()
        offs = np.array(eng.offsets)
        leaves = sorted(set(int(np.searchsorted(offs, b, side="right") - 1) for b in bad[:100000]))
        print("  non-finite gradient leaves:", [eng.paths[j] for j in leaves[:12]], "of", len(leaves))
