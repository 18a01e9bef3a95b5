"""Experiment: throughput of consecutive independent 64-window steps when they alternate over S streams (one handle =
one workspace per stream), with the streams' phases offset so that one step's CNN overlaps another step's transformer.
python tools/staggered_streams_experiment.py"""
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
R = 4
audio = [torch.tensor(synth.make_windows_fast(64, 1234 + r), device="cuda:0") for r in range(R)]


def fwd(e, x, o, s):
    rc = e.L.a2m_forward(e.h, x.data_ptr(), x.shape[0], cos.data_ptr(), sin.data_ptr(), cos.shape[0], o[0].data_ptr(), o[1].data_ptr(), None, 0,
                         C.c_void_p(s.cuda_stream))
    assert rc == 0, rc


def run(S, B=64, iters=40, stagger=True):
    engs = []
    for _ in range(S):
        e = M._Engine(0, model.precision)
        e.load(model)
        engs.append(e)
    streams = [torch.cuda.Stream() for _ in range(S)]
    outs = [(torch.empty(B, 250, 90, device="cuda"), torch.empty(B, 250, 90, device="cuda")) for _ in range(S)]
    small = [(torch.empty(B, 250, 90, device="cuda"), torch.empty(B, 250, 90, device="cuda")) for _ in range(S)]
    for i in range(3 * S):
        fwd(engs[i % S], audio[i % R][:B], outs[i % S], streams[i % S])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    if stagger:      # offset stream k by ~k/S of a step with a partial-size forward that is not counted
        for k in range(1, S):
            nb = max(1, (B * k) // S)
            fwd(engs[k], audio[0][:nb], (small[k][0][:nb], small[k][1][:nb]), streams[k])
    for i in range(iters * S):
        fwd(engs[i % S], audio[i % R][:B], outs[i % S], streams[i % S])
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (iters * S)
    print(f"{S} stream(s) x {B} windows, stagger={stagger}: {ms:.3f} ms per step -> {B * 5 / ms * 1e3:.0f} audio-s/s", flush=True)




def run_prio(prios, B=64, iters=40):
    """ONE handle, len(prios) workspaces, streams with the given CUDA priorities (-1 = high, 0 = default)."""
    S = len(prios)
    e = M._Engine(0, model.precision)
    e.load(model)
    need = int(e.L.a2m_workspace_bytes(e.h, B, 0))
    raw = [torch.zeros(need + 1024, dtype=torch.uint8, device="cuda") for _ in range(S)]
    ws = [(t.data_ptr() + 1023) & ~1023 for t in raw]
    streams = [torch.cuda.Stream(priority=p) for p in prios]
    outs = [(torch.empty(B, 250, 90, device="cuda"), torch.empty(B, 250, 90, device="cuda")) for _ in range(S)]

    def f(i):
        k = i % S
        x = audio[i % R][:B]
        rc = e.L.a2m_forward(e.h, x.data_ptr(), B, cos.data_ptr(), sin.data_ptr(), cos.shape[0], outs[k][0].data_ptr(), outs[k][1].data_ptr(),
                             C.c_void_p(ws[k]), need, C.c_void_p(streams[k].cuda_stream))
        assert rc == 0, rc
    for i in range(3 * S):
        f(i)
    torch.cuda.synchronize()
    n = iters * S
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s_ in streams:
        s_.wait_event(e0)
    for i in range(n):
        f(i)
    for s_ in streams:
        main.wait_stream(s_)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"priorities {prios}: {ms:.3f} ms per step -> {B * 5 / ms * 1e3:.0f} audio-s/s", flush=True)


for pr in ([0, 0], [-1, 0], [-1, -1], [-1, 0, 0], [-2, -1, 0], [0, 0, 0, 0]):
    run_prio(pr)
sys.exit(0)
run(1)
run(2, stagger=False)
run(2)
run(3)
run(2, B=32)
run(4, B=32)
run(1, B=128)


def run_one_handle(S, B=64, iters=40):
    """Same, but ONE handle with S caller-provided workspaces (what model.predict_many does)."""
    e = M._Engine(0, model.precision)
    e.load(model)
    need = int(e.L.a2m_workspace_bytes(e.h, B, 0))
    raw = [torch.zeros(need + 1024, dtype=torch.uint8, device="cuda") for _ in range(S)]
    ws = [(t.data_ptr() + 1023) & ~1023 for t in raw]
    streams = [torch.cuda.Stream() for _ in range(S)]
    outs = [(torch.empty(B, 250, 90, device="cuda"), torch.empty(B, 250, 90, device="cuda")) for _ in range(S)]

    def f(i):
        k = i % S
        x = audio[i % R][:B]
        rc = e.L.a2m_forward(e.h, x.data_ptr(), B, cos.data_ptr(), sin.data_ptr(), cos.shape[0], outs[k][0].data_ptr(), outs[k][1].data_ptr(),
                             C.c_void_p(ws[k]), need, C.c_void_p(streams[k].cuda_stream))
        assert rc == 0, rc
    for i in range(3 * S):
        f(i)
    torch.cuda.synchronize()
    for n in (iters * S, 20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record(main)
        for s in streams:
            s.wait_event(e0)
        for i in range(n):
            f(i)
        for s in streams:
            main.wait_stream(s)
        e1.record(main)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"ONE handle, {S} workspaces x {B} windows, {n} steps: {ms:.3f} ms per step -> {B * 5 / ms * 1e3:.0f} audio-s/s", flush=True)


run_one_handle(2)
run_one_handle(1)
parts = [audio[i % R] for i in range(20)]
for _ in range(2):
    model.predict_many(None, parts[:4], rope)
torch.cuda.synchronize()
for n in (20, 80):
    pp = [audio[i % R] for i in range(n)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.predict_many(None, pp, rope)
    e1.record()
    torch.cuda.synchronize()
    print(f"model.predict_many, {n} steps: {e0.elapsed_time(e1) / n:.3f} ms per step", flush=True)
