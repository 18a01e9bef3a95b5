"""Splits the device eventizer's time (config 5) into its two kernels and the host-side merge: python tools/extract_timing.py"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_to_midi_b200 as A
from audio_to_midi_b200 import infer as I
from oracle import synth

model = A.OutputSequenceGenerator(A.model_config, key=1234)
base = synth.make_clip(30.0, 1243)
rng = np.random.Generator(np.random.PCG64(1244))
clip = np.concatenate([base * np.float32(g) for g in rng.uniform(0.7, 1.3, size=20)], axis=1).astype(np.float32)
dev = torch.device("cuda:0")
rope = A.precompute_frequencies(64, 300)
w = I.prepare_windows_device(model, torch.tensor(clip).to(dev), 0.5)
out = model.predict_many(None, [w[i:i + 64] for i in range(0, w.shape[0], 64)], rope)
st = I.stitch_probs_device(model, torch.cat([p for _, p in out]), 0.5, 0.02)
eng = model._engine(0)
F, K = st.shape
cap = 65536
ev = torch.empty(cap, dtype=torch.int64, device=dev); cnt = torch.empty(1, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream()
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    eng.L.a2m_extract_events_dev(eng.h, st.data_ptr(), F, K, ev.data_ptr(), cap, cnt.data_ptr(), C.c_void_p(s.cuda_stream))
    e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter(); evs = I.extract_events_device(model, st); t1 = time.perf_counter()
    print(f"frames {F}: kernels {e0.elapsed_time(e1):.3f} ms; extract_events_device end to end {1e3 * (t1 - t0):.3f} ms; {len(evs)} events")
