"""Per-launch CUDA-event profile of the forward plan (a2m_profile_steps).  usage: python tools/profile_steps.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_to_midi_b200 as A  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = A.OutputSequenceGenerator(A.model_config, key=1234)
prof = model.profile_steps(B, repeats=10, device=0)
tot = sum(p[1] for p in prof)
print(f"B={B} steps={len(prof)} sum={tot:.3f} ms")
last = None
run = 0
for i, (k, ms, fl, by) in enumerate(prof):
    tf = fl / (ms / 1e3) / 1e12 if fl else 0.0
    gb = by / (ms / 1e3) / 1e9
    print(f"{i:4d} {k:26s} {ms * 1e3:8.1f} us  {tf:7.1f} TF/s  {gb:7.0f} GB/s  flops={fl:.3g} bytes={by:.3g}")
