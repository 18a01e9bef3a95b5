"""Phase timeline of ffn_fused_kernel inside a real forward (needs a build with A2M_EXTRA_NVCC_FLAGS=-DA2M_FFN_TIMING).
usage: A2M_EXTRA_NVCC_FLAGS=-DA2M_FFN_TIMING python audio-to-midi_b200/build.py; python tools/ffn_timeline.py [B]"""
import ctypes as C
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
for _ in range(4):
    model.predict(None, audio, rope)
torch.cuda.synchronize()
buf = (C.c_longlong * 128)()
eng.L.a2m_debug_read_timing.argtypes = [C.c_void_p, C.c_int32]
n = eng.L.a2m_debug_read_timing(buf, 128)
if n < 0:
    raise SystemExit("library was built without -DA2M_FFN_TIMING")
t = list(buf)
t0 = t[0]
rel = lambda i: t[i] - t0
print(f"setup done {rel(1)}  pdl_wait passed {rel(2)}  LN done {rel(3)}  MMA sees A {rel(4)}")
for c in range(8):
    m = [rel(8 + c * 4 + j) for j in range(4)]
    e = [rel(48 + c * 3 + j) for j in range(3)]
    print(f"chunk {c}: MMA d1free {m[0]:6d} mma1 issued {m[1]:6d} hfull seen {m[2]:6d} w2 ready {m[3]:6d} | "
          f"epi d1full {e[0]:6d} tmem read {e[1]:6d} h written {e[2]:6d}")
print(f"done seen {rel(5)}  staged {rel(6)}  end {rel(7)}   (cycles; 1.965 GHz -> {rel(7) / 1965:.1f} us)")

t0 = t[80]
rel = lambda i: t[i] - t0
print(f"qkv: setup done {rel(81)}  LN done {rel(82)}  MMA sees A {rel(83)}")
for n in range(3):
    print(f"chunk {n}: MMA dfree {rel(84 + n * 6):6d}  weights ready kb0..3 {[rel(84 + n * 6 + 1 + k) for k in range(4)]} | epi dfull {rel(104 + 2 * n):6d} stored {rel(105 + 2 * n):6d}")

t0 = t[112]
rel = lambda i: t[i] - t0
print(f"block_fused<128>: pdl_wait passed {rel(113)}  dwconv+LN done {rel(114)}  synced {rel(115)}  D1 ready {rel(116)}  "
      f"gelu0 done {rel(117)} mma2_0 issue {rel(118)}  gelu1 done {rel(119)} mma2_1 issue {rel(120)}  D2 ready {rel(121)}  staged {rel(122)}  end {rel(123)}")

for name, b in (("attn_global", 64), ("attn_local", 72)):
    t0 = t[b]
    print(f"{name}: pdl_wait passed {t[b+1]-t0}  loaded {t[b+2]-t0}  S ready {t[b+3]-t0}  softmax done {t[b+4]-t0}  O ready {t[b+5]-t0}  stored {t[b+6]-t0}")

t0 = t[96]
print("block_mid2<32>: tile loop entered %d  x tile in smem %d  dwconv+LN+A1 written %d  D1 ready %d  GELU+A2 written %d  D2 ready %d  tile 0 done %d  tile 1 done %d"
      % tuple(t[i] - t0 for i in range(97, 105)))

t0 = t[80]
print("qkv (pair or single, whichever ran): setup+cluster sync %d  LN done %d  dfull chunks %s  epilogue done %d  after final cluster sync %d"
      % (t[81] - t0, t[82] - t0, [t[104 + 2 * n] - t0 for n in range(3)], t[110] - t0, t[111] - t0))
