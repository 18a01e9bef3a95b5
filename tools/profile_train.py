"""Per-launch timings of the training plans at batch B (GPU): python tools/profile_train.py 64"""
import sys, time
import numpy as np, torch
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
    eng.training_step(audio, labels, rope, cfg, 1e-4)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
n = 5
tot = np.zeros(4)
for _ in range(n):
    eng.zero_grad()
    ev[0].record()
    cos, sin = eng._rope_tensors(rope)
    eng.L.a2m_forward_train(eng.h, audio.data_ptr(), B, cos.data_ptr(), sin.data_ptr(), 300, None, None, eng._stream())
    ev[1].record()
    eng.L.a2m_backward(eng.h, labels.data_ptr(), 1.0, eng.grads.data_ptr(), eng.loss.data_ptr(), eng._stream())
    ev[2].record()
    eng.optimizer_step(1e-4, cfg)
    ev[3].record()
    torch.cuda.synchronize()
    tot += np.array([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3])])
tot /= n
print(f"B={B} fwd {tot[0]:.3f} ms  bwd {tot[1]:.3f} ms  adamw+repack {tot[2]:.3f} ms  total {tot[3]:.3f} ms  -> {B / tot[3] * 1e3:.0f} samples/s  launches {eng.launch_count()}")
eng.forward_backward(audio, labels, rope)
eng.L.a2m_forward_train(eng.h, audio.data_ptr(), B, cos.data_ptr(), sin.data_ptr(), 300, None, None, eng._stream())
for which, name in ((0, "forward"), (1, "backward")):
    prof = eng.profile_steps(which, 3)
    fam = {}
    for k, ms, fl, by in prof:
        f = fam.setdefault(k, [0.0, 0, 0.0, 0.0]); f[0] += ms; f[1] += 1; f[2] += fl; f[3] += by
    print(f"--- {name}: {len(prof)} steps, sum {sum(p[1] for p in prof):.3f} ms")
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:30s} {f[0]:8.3f} ms  n={f[1]:4d}  {f[2] / f[0] / 1e9 if f[0] else 0:8.1f} TF/s  {f[3] / f[0] / 1e6 if f[0] else 0:8.0f} GB/s")
    if len(sys.argv) > 2:
        for i, (k, ms, fl, by) in enumerate(prof):
            print(f"   {i:4d} {k:30s} {ms * 1e3:8.1f} us  {fl / ms / 1e9 if ms else 0:8.1f} TF/s {by / ms / 1e6 if ms else 0:8.0f} GB/s")
