#!/usr/bin/env python
"""Benchmark of the hot path: batched model forward on synthetic windows (BASELINE.json configs[1]), plus the other
configurations north_star names (train step weak / strong scaling, long-clip transcription, validation pass).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One process per GPU (the driver launches N>1 through torch.distributed.run).  A "step" is one batched forward of B = 64
windows of 5 s per GPU (weak scaling: windows are independent, no collective).  Prints ONE JSON line (rank 0).
See DESIGN.md "Measurement" for how each field is obtained.
"""
from __future__ import annotations

import os

# torchrun exports OMP_NUM_THREADS=1 to every rank.  The CPU arms of this file (--impl reference; cpu_baseline at N = 1) are
# supposed to use all host cores, and the thread count must be in the environment BEFORE numpy / torch load their OpenMP and
# MKL runtimes (round 1 printed a 100x too slow CPU baseline under torchrun for exactly this reason).
_CORES = len(os.sched_getaffinity(0))
if int(os.environ.get("RANK", "0")) == 0:
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(_CORES)

import argparse  # noqa: E402
import json  # noqa: E402
import statistics  # noqa: E402
import subprocess  # noqa: E402
import sys  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WINDOW_S = 5.0
FLOPS_PER_WINDOW = 2 * 3_618_361_856  # de-duplicated forward MACs x 2
TRAIN_FLOPS_PER_SAMPLE = 3 * FLOPS_PER_WINDOW   # fwd + dgrad + wgrad, no recompute counted
SEED = 1234
DROPOUT = 0.1   # transformer_dropout_rate (model.py:30), applied as in train.py:56-58 (enable_dropout=True)
PORT_NOTE = "PyTorch-CPU fp32 restatement (oracle/model_torch.py); the JAX reference is not installable (SURVEY F1)"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "tensor_burst": p["bf16_tflops"], "tensor_sustained": p["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons of one GPU during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ CPU arms (oracle port)
def cpu_forward_rate(n_windows: int, iters: int, threads: int, warmup: int = 1):
    """Times the oracle's PyTorch-CPU fp32 restatement.  Returns (audio-seconds per second, seconds per iteration)."""
    import torch
    from oracle import model_torch as T
    from oracle import params as P
    from oracle import synth
    torch.set_num_threads(threads)
    params = T.to_torch(P.init_params(SEED))
    audio = torch.tensor(synth.make_windows_fast(n_windows, SEED))
    with torch.no_grad():
        for _ in range(max(warmup, 1)):
            T.forward(params, audio)
        times = []
        for _ in range(iters):
            t0 = time.perf_counter()
            T.forward(params, audio)
            times.append(time.perf_counter() - t0)
    best = statistics.median(times)
    return n_windows * WINDOW_S / best, best


def cpu_train_rate(n_windows: int, iters: int, threads: int):
    """Forward + backward (torch autograd) of the oracle's loss_fn (train.py:39-62) on the host cores, on a bounded sample.
    Optimizer time is not included (< 1 % on the CPU).  Returns (samples per second, seconds per iteration)."""
    import torch
    from oracle import model_torch as T
    from oracle import params as P
    from oracle import synth
    torch.set_num_threads(threads)
    tree = T.to_torch(P.init_params(SEED), requires_grad=False)
    leaves = []

    def mark(t):
        if isinstance(t, dict):
            return {k: mark(v) for k, v in t.items()}
        if isinstance(t, list):
            return [mark(v) for v in t]
        if t.dtype.is_floating_point:
            t = t.clone().requires_grad_(True)
            leaves.append(t)
        return t

    tree = mark(tree)
    audio, labels = synth.make_windows(n_windows, SEED, with_labels=True)
    audio, labels = torch.tensor(audio), torch.tensor(labels)
    times = []
    for it in range(iters + 1):
        for t in leaves:
            t.grad = None
        t0 = time.perf_counter()
        loss, _ = T.loss_fn(tree, audio, labels)
        loss.backward()
        if it > 0:   # the first pass warms the thread pool and the allocator
            times.append(time.perf_counter() - t0)
    best = statistics.median(times)
    return n_windows / best, best


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores.  The JAX reference cannot be
    imported (no jax / equinox in the image; /root/reference does not exist on the GPU box), so this times the oracle port --
    on the SAME step as our arm (args.batch windows per GPU per step, same warm-up and step counts), bounded to 128 windows per
    step so that an N = 8 launch still ends within a few minutes (the metric, audio-seconds per second, does not depend on it)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import model_torch as T
    from oracle import params as P
    from oracle import synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = _CORES
    torch.set_num_threads(cores)
    sample = min(args.batch * world, 128)
    W = max(args.warmup, 3)
    params = T.to_torch(P.init_params(SEED))
    audio = torch.tensor(synth.make_windows_fast(sample, SEED))
    with torch.no_grad():
        for _ in range(W):
            T.forward(params, audio)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            T.forward(params, audio)
        dt = time.perf_counter() - t0
        # config 1 (BASELINE.json configs[0]): one clip window, batch 1
        one = audio[:1]
        T.forward(params, one)
        t1 = time.perf_counter()
        for _ in range(5):
            T.forward(params, one)
        dt1 = (time.perf_counter() - t1) / 5
    value = sample * WINDOW_S * args.steps / dt
    line = {
        "impl": "reference", "metric": "audio-seconds/sec transcribed (fwd)", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": W, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batched forward of the default model, {args.batch} synthetic 5 s windows per GPU "
                               f"(BASELINE.json configs[1]), random-init weights, windows batch-partitioned across GPUs",
                   "batch_per_gpu": args.batch, "cpu_windows_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} windows per step x {args.steps} steps after {W} warm-ups; {PORT_NOTE}",
                         "config1_batch1": {"value": WINDOW_S / dt1, "unit": "audio-s/s", "ms_per_window": dt1 * 1e3,
                                            "what": "BASELINE.json configs[0]: one 5 s window, batch 1, CPU"}},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm: training
class Ctx:
    pass


def _max_over_ranks(ctx, x: float) -> float:
    import torch
    t = torch.tensor([x], device=ctx.dev, dtype=torch.float64)
    if ctx.dist is not None:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(ctx):
    import torch
    if ctx.dist is not None:
        ctx.dist.barrier()
    torch.cuda.synchronize()


def measure_train(ctx, args, eng, B, K, label, e2e=True):
    """One training configuration: B windows per GPU per step; forward with tape, backward, gradient all-reduce (a2m_allreduce_grads:
    NCCL inside the library, bucket 0 under the CNN backward), AdamW + clip, weight re-pack (train.py:259-332)."""
    import torch
    from audio_to_midi_b200 import train as T
    A, synth, rank, world = ctx.A, ctx.synth, ctx.rank, ctx.world
    cfg = T.OptimizerConfig()
    sched = T.create_learning_rate_schedule(cfg.base_learning_rate, cfg.warmup_steps, cfg.num_steps)
    rope = ctx.rope
    R = 3
    host = [synth.make_windows_fast(B, SEED + 31 * (rank * R + r)) for r in range(R)]
    rng = np.random.Generator(np.random.PCG64(SEED + rank))
    host_y = [np.clip((rng.random((B, 250, 90)) < 0.02).astype(np.float32), 0.005, 0.995) for _ in range(R)]
    dev_x = [torch.tensor(h, device=ctx.dev) for h in host]
    dev_y = [torch.tensor(h, device=ctx.dev) for h in host_y]
    W = max(args.warmup, 3)
    step0 = eng.step_count + 1000      # past the zero-lr first update of the warm-up schedule
    for i in range(W):
        eng.training_step(dev_x[i % R], dev_y[i % R], rope, cfg, sched(step0 + i), dropout_rate=DROPOUT, key=SEED)
    _barrier(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        eng.training_step(dev_x[i % R], dev_y[i % R], rope, cfg, sched(step0 + W + i), dropout_rate=DROPOUT, key=SEED)
    e1.record()
    _barrier(ctx)
    ms = _max_over_ranks(ctx, e0.elapsed_time(e1))
    loss_last = float(eng.loss.item())
    # phase breakdown (this rank's events; separate untimed pass).  "allreduce" is the EXPOSED time: from the end of the backward
    # on the compute stream to the point where that stream holds the fully reduced gradients.
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    acc = np.zeros(4)
    cos, sin = eng._rope_tensors(rope)
    for i in range(3):
        eng.zero_grad()
        eng.set_dropout(DROPOUT, SEED + i)
        ev[0].record()
        eng.L.a2m_forward_train(eng.h, dev_x[i % R].data_ptr(), B, cos.data_ptr(), sin.data_ptr(), 300, None, None, eng._stream())
        ev[1].record()
        eng.L.a2m_backward(eng.h, dev_y[i % R].data_ptr(), 1.0, eng.grads.data_ptr(), eng.loss.data_ptr(), eng._stream())
        ev[2].record()
        eng.allreduce_grads()
        ev[3].record()
        eng.optimizer_step(sched(step0), cfg)
        ev[4].record()
        torch.cuda.synchronize()
        acc += np.array([ev[j].elapsed_time(ev[j + 1]) for j in range(4)])
    acc /= 3
    out = {
        "metric": "train samples/sec", "value": world * B * K / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms / K, "steps": K,
        "warmup": W, "batch_per_gpu": B, "global_batch": B * world, "scaling": label,
        "step": "forward with tape + backward + gradient all-reduce (a2m_allreduce_grads: NCCL, two buckets, the first under the CNN "
                "backward) + AdamW/clip + weight re-pack (train.py:259-332)",
        "dtype": "bf16 operands, fp32 accumulate / master weights / optimizer state (reference: fp16 forward+backward, fp32 master)",
        "dropout": DROPOUT, "allreduce_backend": "a2m (in-library NCCL)" if eng.comm_ready else ("torch.distributed" if world > 1 else "none"),
        "breakdown_ms": {"forward": round(float(acc[0]), 3), "backward": round(float(acc[1]), 3),
                         "allreduce_exposed": round(float(acc[2]), 3), "adamw_repack": round(float(acc[3]), 3)},
        "tflops": world * B * TRAIN_FLOPS_PER_SAMPLE / (ms / K / 1e3) / 1e12,
        "frac_of_tensor_peak": B * TRAIN_FLOPS_PER_SAMPLE / (ms / K / 1e3) / 1e12 / _peaks()["tensor_sustained"],
        "gpu_launches": eng.launch_count() * K, "loss_after_timed": loss_last, "allreduce_bytes": eng.n_params * 4,
    }
    if e2e:
        # end to end: pinned host audio + labels copied in every step, loss + validity read back every step
        pin_x = [torch.tensor(h).pin_memory() for h in host]
        pin_y = [torch.tensor(h).pin_memory() for h in host_y]
        batches = [(pin_x[i % R], pin_y[i % R]) for i in range(K)]
        eng.train_pipelined(batches[:2], rope, cfg, sched, first_step=step0, dropout_rate=DROPOUT, key=SEED)   # warm the copy path
        _barrier(ctx)
        t0 = time.perf_counter()
        losses = eng.train_pipelined(batches, rope, cfg, sched, first_step=step0, dropout_rate=DROPOUT, key=SEED)
        torch.cuda.synchronize()
        dt = _max_over_ranks(ctx, time.perf_counter() - t0)
        out["e2e"] = {"value": world * B * K / dt, "unit": "samples/s", "h2d_bytes_per_step": B * (2 * 80000 + 250 * 90) * 4,
                      "d2h_bytes_per_step": 8,
                      "api": "TrainEngine.train_pipelined (next batch's H2D on a copy stream; loss and grads_valid read back every step, "
                             "consumed one step later)"}
        out["last_loss"] = losses[-1]
    return out


def train_roofline(eng):
    """Dominant kernel family of the training step (per-launch CUDA events of both plans), against the tensor roof."""
    peaks = _peaks()
    fam = {}
    for which in (0, 1):
        for k, pms, fl, by in eng.profile_steps(which, repeats=3):
            f = fam.setdefault(k, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            f["ms"] += pms; f["flops"] += fl; f["bytes"] += by; f["n"] += 1
    total = sum(f["ms"] for f in fam.values())
    name, f = max(fam.items(), key=lambda kv: kv[1]["ms"])
    ach = f["flops"] / (f["ms"] / 1e3) / 1e12
    traffic = None
    try:
        import glob
        tf = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_train_traffic.json")))[-1]
        with open(tf) as fh:
            traffic = json.load(fh).get(name)
    except Exception:
        pass
    return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peaks["tensor_sustained"], "unit": "TFLOP/s",
            "frac": ach / peaks["tensor_sustained"], "traffic": traffic, "launches_per_step": f["n"], "kernel_ms_per_step": f["ms"],
            "share_of_serialised_step": f["ms"] / total,
            "hbm": {"achieved": f["bytes"] / (f["ms"] / 1e3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": f["bytes"] / (f["ms"] / 1e3) / 1e9 / peaks["hbm"]},
            "note": "per-launch times are serialised (side-stream wgrad kernels overlap the main chain in the real step)",
            "families_ms": {k: round(v["ms"], 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}}


def dp_selfcheck(ctx, eng_small):
    """Data-parallel correctness on the GPUs: the gradients after the all-reduce over `world` ranks x b windows
    equal the 1-rank gradients of the concatenated batch.  Dropout off (masks are indexed by the local sample position)."""
    import torch
    A, synth, rank, world = ctx.A, ctx.synth, ctx.rank, ctx.world
    b = 2
    audio, labels = synth.make_windows(b * world, SEED + 5, with_labels=True)      # identical on every rank (same seed)
    x = torch.tensor(audio, device=ctx.dev)
    y = torch.tensor(labels, device=ctx.dev)
    out = {}
    for backend in (["a2m", "torch"] if eng_small.comm_ready else ["torch"]):
        eng_small.zero_grad()
        eng_small.set_dropout(0.0, 0)
        eng_small.forward_backward(x[rank * b:(rank + 1) * b], y[rank * b:(rank + 1) * b], ctx.rope)
        eng_small.allreduce_grads(backend=backend)
        torch.cuda.synchronize()
        out[backend] = (eng_small.grads.clone(), float(eng_small.loss.item()))
    eng_small.zero_grad()
    eng_small.forward_backward(x, y, ctx.rope)
    torch.cuda.synchronize()
    g1, l1 = eng_small.grads.clone(), float(eng_small.loss.item())
    res = {"ranks": world, "windows_per_rank": b}
    for backend, (g, l) in out.items():
        rel = float(((g - g1).norm() / g1.norm()).item())
        res[backend] = {"grad_rel_l2_vs_one_rank": rel, "loss_rel": abs(l - l1) / abs(l1), "ok": bool(rel < 1e-3)}
    return res


# ------------------------------------------------------------------------------------------------ our arm: configs 5 and 3
def measure_clip(ctx, model):
    """BASELINE config 5: a 10-minute clip (9.6 M samples per channel) -> device normalise + slice (134 windows at 0.5 s overlap) ->
    this rank's block through the batched forward -> rank-ordered all_gather of the probabilities -> rank 0 stitches (30 175 frames)
    and eventizes.  Host clip in page-locked memory; the timed region starts at the H2D copy and ends when rank 0 holds the events."""
    import torch
    from audio_to_midi_b200 import infer as I
    base = ctx.synth.make_clip(30.0, SEED + 9)
    rng = np.random.Generator(np.random.PCG64(SEED + 10))
    clip = np.concatenate([base * np.float32(g) for g in rng.uniform(0.7, 1.3, size=20)], axis=1).astype(np.float32)
    assert clip.shape == (2, 9_600_000)
    pin = torch.tensor(clip).pin_memory()
    times, n_events, frames, n_windows = [], 0, 0, 0
    for it in range(4):
        _barrier(ctx)
        t0 = time.perf_counter()
        dev_clip = pin.to(ctx.dev, non_blocking=True)
        events, _st, _pr = I.transcribe_clip(model, dev_clip, overlap=0.5, want_arrays=False)
        torch.cuda.synchronize()
        dt = _max_over_ranks(ctx, time.perf_counter() - t0)
        if it > 0:
            times.append(dt)
        if ctx.rank == 0:
            n_windows = int(model._engine(ctx.dev.index).L.a2m_window_count(clip.shape[1], 0.5))
            n_events, frames = len(events), n_windows * 250 - 25 * (n_windows - 1)
    best = statistics.median(times)
    return {"metric": "clip audio-seconds/sec transcribed end to end", "value": 600.0 / best, "unit": "audio-s/s", "seconds_per_clip": best,
            "clip_seconds": 600.0, "windows": n_windows, "windows_per_gpu": -(-n_windows // ctx.world), "stitched_frames": frames, "events": n_events,
            "h2d_bytes": int(clip.nbytes) * ctx.world, "d2h": "the event list only (one 64-bit word per event): stitch_probs and extract_events run on the device",
            "gather": "all_gather of [windows/N, 250, 90] fp32 blocks (NCCL), rank 0 stitches + eventizes on its GPU" if ctx.world > 1 else "none (one rank)",
            "what": "BASELINE.json configs[4]; infer.transcribe_clip; median of 3 after 1 warm-up; every rank uploads and normalises the whole clip "
                    "(the loudness statistics need all of it), then forwards only its block of windows"}


def measure_eval(ctx, model):
    """BASELINE config 3: validation loss / hit-rate over 512 annotated windows, windows batch-partitioned over the ranks; per-window
    BCE and event metrics on the device, [512, 6] floats gathered in rank order."""
    import torch
    from audio_to_midi_b200 import infer as I
    a, y = ctx.synth.make_windows(32, SEED + 3, with_labels=True)
    audio = np.tile(a, (16, 1, 1))
    labels = np.tile(y, (16, 1, 1))
    times = []
    for it in range(3):
        _barrier(ctx)
        t0 = time.perf_counter()
        lo, hi, losses, details = I.compute_testset_loss(model, audio, labels, max_batch=64)
        torch.cuda.synchronize()
        dt = _max_over_ranks(ctx, time.perf_counter() - t0)
        if it > 0:
            times.append(dt)
    best = statistics.median(times)
    hit = float(np.mean([d["hit_rate"] for d in details]))
    # the same with the annotated set already resident on the device (this rank's block is all it touches)
    dev_a, dev_y = torch.tensor(audio[lo:hi], device=ctx.dev), torch.tensor(labels[lo:hi], device=ctx.dev)
    dtimes = []
    for it in range(3):
        _barrier(ctx)
        t0 = time.perf_counter()
        _lo, _hi, dl, _dd = I.compute_testset_loss(model, dev_a, dev_y, rank=0, world_size=1, max_batch=64)
        torch.cuda.synchronize()
        dt = _max_over_ranks(ctx, time.perf_counter() - t0)
        if it > 0:
            dtimes.append(dt)
    dbest = statistics.median(dtimes)
    assert np.allclose(dl, losses[lo:hi] if len(losses) == 512 else losses, rtol=1e-5)
    return {"metric": "validation windows/sec", "value": 512 / best, "unit": "windows/s", "seconds": best, "windows": 512,
            "device_resident_value": 512 / dbest, "device_resident_seconds": dbest,
            "mean_loss": float(np.mean(losses)), "mean_hit_rate": hit, "results_gathered": int(len(losses)),
            "what": "BASELINE.json configs[2]; infer.compute_testset_loss from pageable host arrays (H2D inside; `value`) and with the set "
                    "resident on the device (`device_resident_value`: each rank's block through predict_many, no gather), a2m_window_losses + "
                    "a2m_event_metrics on the device; random-init weights, so the hit rate itself is meaningless"}


# ------------------------------------------------------------------------------------------------ our arm: forward
def run_ours(args):
    import torch
    import audio_to_midi_b200 as A
    from audio_to_midi_b200 import hostbind
    from audio_to_midi_b200 import train as T
    from oracle import synth  # synthetic inputs only (seeded generator); not on the measured path

    ctx = Ctx()
    ctx.A, ctx.synth = A, synth
    ctx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    placement = hostbind.bind_to_gpu_numa(local)       # before any page-locked allocation
    ctx.dev = dev = torch.device("cuda", local)

    # ---- CPU baselines: N = 1 only, BEFORE any process group exists (no rank ever waits in a collective for CPU work)
    cpu_fwd = cpu_train = None
    if world == 1 and not args.no_cpu:
        cores = _CORES
        cpu_n, cpu_iters = 64, 3
        rate, sec = cpu_forward_rate(cpu_n, cpu_iters, cores)
        rate1, sec1 = cpu_forward_rate(1, 5, cores)
        cpu_fwd = {"value": rate, "unit": "audio-s/s", "cores": cores, "kind": "port",
                   "sample": f"the same {cpu_n}-window step x {cpu_iters} iterations ({sec:.2f} s each); {PORT_NOTE}",
                   "config1_batch1": {"value": rate1, "unit": "audio-s/s", "ms_per_window": sec1 * 1e3,
                                      "what": "BASELINE.json configs[0]: one 5 s window, batch 1, CPU"}}
        trate, tsec = cpu_train_rate(8, 2, cores)
        cpu_train = {"value": trate, "unit": "samples/s", "cores": cores, "kind": "port",
                     "sample": f"forward + backward of 8 windows x 2 iterations ({tsec:.2f} s each), torch autograd of the oracle's loss_fn; {PORT_NOTE}"}

    ctx.dist = dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        ctx.dist = dist

    B = args.batch
    model = A.OutputSequenceGenerator(A.model_config, key=SEED)      # random-init weights of the reference architecture
    ctx.rope = rope = A.precompute_frequencies(A.model_config["attention_size"], 300)
    # inputs: R distinct batches so that consecutive steps read different audio (R * B * 640 KB > 126 MB L2)
    R = max(2, -(-140_000_000 // (B * 640_000)))
    host_batches = [synth.make_windows_fast(B, SEED + 17 * (rank * R + r)) for r in range(R)]
    dev_batches = [torch.tensor(h, device=dev) for h in host_batches]
    predict = A.vmap(model.predict, in_axes=(None, 0, None))
    W = max(args.warmup, 3)

    # ---- warm-up (also builds plan + CUDA graph)
    for i in range(W):
        predict(None, dev_batches[i % R], rope)
    torch.cuda.synchronize()
    launches_per_step = model.last_launch_count(local)

    # ---- device-resident timing: exactly K steps between two events on the launching stream.  The K independent batches go
    # through model.predict_many: two lanes (streams + workspaces), so consecutive steps overlap; the one-after-the-other time of
    # a single step is measured next to it (config.serial_ms_per_step).
    steps_in = [dev_batches[i % R] for i in range(args.steps)]
    for _ in range(2):
        model.predict_many(None, steps_in, rope)       # also lets the caching allocator keep the output block of a K-step call
    torch.cuda.synchronize()
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record()
    for i in range(args.steps):
        predict(None, dev_batches[i % R], rope)
    es1.record()
    torch.cuda.synchronize()
    serial_ms = _max_over_ranks(ctx, es0.elapsed_time(es1)) / args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(ctx)
    with ClockSampler(local) as clocks:
        e0.record()
        model.predict_many(None, steps_in, rope)
        e1.record()
        _barrier(ctx)
        ms = e0.elapsed_time(e1)
        # keep the sampler alive long enough for at least a few samples under load
        t_end = time.perf_counter() + max(0.0, 1.0 - ms / 1e3)
        while time.perf_counter() < t_end:
            predict(None, dev_batches[0], rope)
        torch.cuda.synchronize()
    ms = _max_over_ranks(ctx, ms)
    value = world * B * WINDOW_S * args.steps / (ms / 1e3)

    # ---- end to end through the public API with HOST buffers: every step's H2D (from page-locked memory) and D2H are inside the
    # timed region.  model.predict_pipelined keeps two batches in flight (copy/compute overlap).  Three byte budgets:
    #   headline  f16 audio in (lossless: the loader rounds to f16, python.rs:235-264), probabilities only, fp32 (what infer.py:41 keeps)
    #   compact   f16 in, probabilities only, f16 out
    #   full_f32  fp32 in, logits + probabilities fp32 out (round 1's path)
    def e2e_run(pinned, **kw):
        for _ in model.predict_pipelined((pinned[i % R] for i in range(A.model.HOST_SLOTS + 3)), rope, **kw):
            pass                                           # every slot's device buffers and the whole page-locked output ring exist now
        _barrier(ctx)
        t0 = time.perf_counter()
        n_done, last = 0, None
        for _lg, pr in model.predict_pipelined((pinned[i % R] for i in range(args.steps)), rope, **kw):
            n_done += 1
            last = pr
        dt = time.perf_counter() - t0
        assert n_done == args.steps and float(last[0, 0, 0]) == float(last[0, 0, 0])
        return world * B * WINDOW_S * args.steps / _max_over_ranks(ctx, dt)

    pinned32, pinned16 = [], []
    for hb in host_batches:
        p32 = A.pinned_empty(hb.shape)
        p32[...] = hb
        p16 = A.pinned_empty(hb.shape, np.float16)
        p16[...] = hb                                   # exact: the synthetic audio is f16-rounded like the loader's
        pinned32.append(p32)
        pinned16.append(p16)
    e2e = e2e_run(pinned16, want_logits=False)
    e2e_compact = e2e_run(pinned16, want_logits=False, probs_dtype=np.float16)
    e2e_full = e2e_run(pinned32)
    # the plain synchronous call (numpy in, numpy out, pageable memory), for comparison
    predict(None, host_batches[0], rope)
    t0 = time.perf_counter()
    for i in range(min(args.steps, 5)):
        predict(None, host_batches[i % R], rope)
    e2e_sync = B * WINDOW_S * min(args.steps, 5) / (time.perf_counter() - t0)
    del pinned32, pinned16

    # ---- roofline of the dominant kernel, from per-launch CUDA-event timings of the same plan (rank 0)
    roof = None
    if rank == 0:
        peaks = _peaks()
        # three passes, per-launch median: one pass (5 repeats per launch) occasionally carries an outlier of 5-10 % on a family
        passes = [model.profile_steps(B, repeats=5, device=local) for _ in range(3)]
        prof = [(p0[0], statistics.median(q[i][1] for q in passes), p0[2], p0[3]) for i, p0 in enumerate(passes[0])]
        fam = {}
        for k, pms, fl, by in prof:
            f = fam.setdefault(k, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            f["ms"] += pms; f["flops"] += fl; f["bytes"] += by; f["n"] += 1
        total_ms = sum(f["ms"] for f in fam.values())
        name, f = max(fam.items(), key=lambda kv: kv[1]["ms"])
        # the path is held against the TENSOR roof (intensity 8 800 FLOP/B against a balance of ~210); the kernel's
        # algorithmic bytes against the HBM peak are kept as a second figure
        achieved = f["flops"] / (f["ms"] / 1e3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["tensor_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tensor_sustained"], "traffic": None,
                "hbm": {"achieved": f["bytes"] / (f["ms"] / 1e3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": f["bytes"] / (f["ms"] / 1e3) / 1e9 / peaks["hbm"],
                        "note": "algorithmic bytes of the kernel / its time; its operands are L2-resident in the steady state"}}
        try:
            import glob
            tf = sorted(f for f in glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")) if "train" not in os.path.basename(f))[-1]
            with open(tf) as fh:
                roof["traffic"] = json.load(fh).get(name)
            roof["traffic_source"] = os.path.relpath(tf, ROOT) + " (ncu dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch)"
        except Exception:
            pass
        roof.update({"kernel": name, "launches_per_step": f["n"], "kernel_ms_per_step": f["ms"],
                     "share_of_step": f["ms"] / total_ms, "peak_source": peaks["src"] + ", sustained bf16 figure",
                     "flops_per_launch_avg": f["flops"] / f["n"], "bytes_per_launch_avg": f["bytes"] / f["n"],
                     "whole_step_tflops": B * FLOPS_PER_WINDOW / (ms / args.steps / 1e3) / 1e12,
                     "whole_step_frac": B * FLOPS_PER_WINDOW / (ms / args.steps / 1e3) / 1e12 / peaks["tensor_sustained"],
                     "families_ms": {k: round(v["ms"], 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}})

    # ---- configs 5 and 3 (inference-side, sharded by windows)
    clip = evalr = None
    if not args.no_extra:
        clip = measure_clip(ctx, model)
        evalr = measure_eval(ctx, model)

    # ---- training step (config 4): weak scaling (64 windows per GPU) and the reference's own global batch of 64 (strong)
    train = train_b64 = None
    if not args.no_train:
        del dev_batches
        torch.cuda.empty_cache()
        tmodel = A.OutputSequenceGenerator(A.model_config, key=SEED)
        eng = T.TrainEngine(tmodel, local)
        eng.set_lr_multipliers(T.layer_lr_multipliers(eng.paths, T.OptimizerConfig().layer_lr_decay))
        if world > 1:
            eng.init_comm()
        train = measure_train(ctx, args, eng, args.train_batch, args.train_steps, "weak")
        if rank == 0:
            train["roofline"] = train_roofline(eng)
        train["cpu_baseline"] = cpu_train
        if world > 1 and 64 % world == 0:
            train_b64 = measure_train(ctx, args, eng, 64 // world, args.train_steps, "strong", e2e=False)
            train_b64["what"] = "the reference's own configuration: global batch 64 (train.py:743-744), 64 / N windows per GPU"
            train_b64["dp_selfcheck"] = dp_selfcheck(ctx, eng)
        elif world == 1:
            train_b64 = {"same_as": "train (one GPU: global batch 64 = 64 windows per GPU)", "value": train["value"], "unit": "samples/s",
                         "ms_per_step": train["ms_per_step"], "global_batch": 64, "scaling": "strong"}
        eng.close()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    line = {
        "metric": "audio-seconds/sec transcribed (fwd)", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": model.precision, "data": "synthetic",
        "config": {"workload": f"batched forward of the default model, {B} synthetic 5 s windows per GPU "
                               f"(BASELINE.json configs[1]), random-init weights, windows batch-partitioned across GPUs",
                   "batch_per_gpu": B, "l2": f"inputs rotated over {R} distinct batches ({R * B * 0.64:.0f} MB > 126 MB L2); "
                                             "weights and activations stay L2-resident as in steady-state serving",
                   "operands": f"{model.precision} tensor-core operands (IEEE binary16 is the inference default: 8x smaller rounding than bf16 "
                               "at the same tcgen05 rate; the reference infers in fp32, infer.py:234)",
                   "accumulate": "fp32", "residual_stream": "fp32", "cuda_graph": True, "host_placement": placement,
                   "lanes": "the K steps run through model.predict_many: two streams x two workspaces, consecutive independent batches overlap",
                   "serial_ms_per_step": serial_ms, "serial_value": world * B * WINDOW_S / (serial_ms / 1e3)},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": B * 2 * 80000 * 2, "d2h_bytes_per_step": B * 250 * 90 * 4,
                "api": "model.predict_pipelined(want_logits=False) over a2m_submit_host_ex: f16 audio from page-locked memory (lossless, the "
                       "loader rounds to f16), fp32 probabilities back (infer.py:41 keeps only those); two batches in flight",
                "compact_value": e2e_compact, "compact_bytes_per_step": [B * 2 * 80000 * 2, B * 250 * 90 * 2],
                "full_f32_value": e2e_full, "full_f32_bytes_per_step": [B * 2 * 80000 * 4, 2 * B * 250 * 90 * 4],
                "sync_call_value": e2e_sync},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roof,
        "cpu_baseline": cpu_fwd if cpu_fwd is not None else {"value": None, "unit": "audio-s/s", "cores": _CORES, "kind": "port",
                                                              "sample": "timed at N = 1 only (and by --impl reference at every N)"},
    }
    for k, v in (("train", train), ("train_b64", train_b64), ("clip", clip), ("eval", evalr)):
        if v is not None:
            line[k] = v
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="windows per GPU per step")
    ap.add_argument("--train-batch", type=int, default=64, help="training windows per GPU per step (weak scaling)")
    ap.add_argument("--train-steps", type=int, default=30)
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurements")
    ap.add_argument("--no-extra", action="store_true", help="skip the clip (config 5) and validation (config 3) measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baselines")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
