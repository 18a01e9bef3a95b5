"""Experiment: does splitting a 64-window step into two 32-window halves on two streams (two handles) raise the forward
throughput by letting different kernels of the two halves overlap?  python tools/two_stream_experiment.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import audio_to_midi_b200 as A  # noqa: E402
from audio_to_midi_b200 import model as M  # noqa: E402
from oracle import synth  # noqa: E402

model = A.OutputSequenceGenerator(A.model_config, key=1234)
rope = A.precompute_frequencies(64, 300)
cos = torch.as_tensor(np.ascontiguousarray(rope.cos_freq, np.float32)).cuda()
sin = torch.as_tensor(np.ascontiguousarray(rope.sin_freq, np.float32)).cuda()
audio = torch.tensor(synth.make_windows_fast(64, 1234), device="cuda:0")


def run(nsplit, iters=30):
    B = 64 // nsplit
    engs = []
    for _ in range(nsplit):
        e = M._Engine(0)
        e.load(model)
        engs.append(e)
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    outs = [(torch.empty(B, 250, 90, device="cuda"), torch.empty(B, 250, 90, device="cuda")) for _ in range(nsplit)]
    xs = [audio[i * B:(i + 1) * B].contiguous() for i in range(nsplit)]

    def step():
        for e, s, o, x in zip(engs, streams, outs, xs):
            rc = e.L.a2m_forward(e.h, x.data_ptr(), B, cos.data_ptr(), sin.data_ptr(), cos.shape[0], o[0].data_ptr(), o[1].data_ptr(), None, 0,
                                 C.c_void_p(s.cuda_stream))
            assert rc == 0, rc
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    for _ in range(iters):
        step()
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"split {nsplit} x {B} windows: {ms:.3f} ms per 64 windows -> {64 * 5 / ms * 1e3:.0f} audio-s/s", flush=True)
    return torch.cat([o[1] for o in outs])


p1 = run(1)
p2 = run(2)
p4 = run(4)
print("max |dprob| split2 vs 1:", float((p1 - p2).abs().max()), "split4:", float((p1 - p4).abs().max()))
