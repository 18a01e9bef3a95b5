"""Host side of the training step: mirror of the reference's train.py hot path over the C ABI.

Reference                                         here
------------------------------------------------  --------------------------------------------------------------
compute_loss (train.py:48-62)                      compute_loss(...) -> ((loss, state), grads)   [value_and_grad]
compute_training_step (train.py:259-332)           TrainEngine.training_step(...)
setup_optimizers (train.py:646-728)                OptimizerConfig + layer_lr_multipliers(...)
create_learning_rate_schedule (train.py:454-466)   create_learning_rate_schedule(...)
batch sharding over devices (train.py:238-244)     one process per GPU; TrainEngine.allreduce_grads() = NCCL all-reduce

Device arrays are torch CUDA tensors (torch is the allocator / stream / NCCL provider only); every kernel is in
libaudio2midi_b200.so.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import numpy as np

from . import _lib
from .model import OutputSequenceGenerator, _Engine, _default_device, model_config
from .rope import RopeFreqs


@dataclass
class OptimizerConfig:  # train.py:691-726, 743-749
    base_learning_rate: float = 1e-4
    layer_lr_decay: float = 0.7
    weight_decay: float = 0.005
    warmup_steps: int = 1000
    num_steps: int = 200_000
    eps: float = 1e-3
    b1: float = 0.9
    b2: float = 0.999
    clip_norm: float = 1.0


def create_learning_rate_schedule(base_learning_rate: float, warmup_steps: int, cosine_decay_steps: int) -> Callable[[int], float]:
    """optax.join_schedules([linear 0 -> base over warmup, cosine_decay(base, steps)], [warmup])  (train.py:454-466)."""
    def schedule(step: int) -> float:
        if step < warmup_steps:
            return base_learning_rate * step / max(warmup_steps, 1)
        t = min(step - warmup_steps, cosine_decay_steps) / max(cosine_decay_steps, 1)
        return base_learning_rate * 0.5 * (1.0 + math.cos(math.pi * t))
    return schedule


def layer_lr_multipliers(paths, layer_lr_decay: float, depths=None) -> np.ndarray:
    """Per-leaf learning-rate multiplier of setup_optimizers (train.py:648-704): leaves under `layers.<stage>.layers.<k>`
    get decay ** (max_depth - depth), depth = sum(depths[:stage]) + k; everything else 1."""
    depths = model_config["depths"] if depths is None else depths
    d = []
    for p in paths:
        parts = p.split(".")
        if parts[0] == "layers":
            d.append(sum(depths[: int(parts[1])]) + int(parts[3]))
        else:
            d.append(None)
    mx = max(x for x in d if x is not None)
    return np.array([1.0 if x is None else layer_lr_decay ** (mx - x) for x in d], np.float32)


def shard_batch(global_batch: int, world_size: int, rank: int):
    """Rows of the global batch owned by `rank`: the batch axis is split evenly over devices (train.py:238-244,
    PartitionSpec("batch")); the global batch must be divisible by the device count, as in the reference."""
    if global_batch % world_size != 0:
        raise ValueError("the batch must be divisible by the number of devices (train.py:744)")
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def allreduce_mean_(tensors):
    """In-place mean over ranks of every tensor (sum all-reduce, then 1/world): what jit does for a batch-sharded mean loss
    (train.py:61-62 under the sharding of train.py:238-244).  NCCL on CUDA tensors, gloo on CPU tensors; no-op for one rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensors
    w = dist.get_world_size()
    for t in tensors:
        dist.all_reduce(t)
        t.mul_(1.0 / w)
    return tensors


class TrainEngine:
    """Master parameters, AdamW state and activation tape on one GPU (a2m_train_init ...)."""

    def __init__(self, model: OutputSequenceGenerator, device: Optional[int] = None):
        import torch
        self.torch = torch
        self.device = _default_device() if device is None else device
        self.eng = _Engine.get(self.device)
        self.L, self.h = self.eng.L, self.eng.h
        leaves = model.tree_leaves_with_path()
        self.paths = [p for p, _ in leaves]
        self.shapes = [tuple(np.shape(a)) for _, a in leaves]
        n = len(leaves)
        table = (_lib.LeafDesc * n)()
        chunks, off, self.offsets = [], 0, []
        for i, (path, arr) in enumerate(leaves):
            a = np.ascontiguousarray(arr, dtype=np.float32)
            table[i].path = path.encode()
            table[i].offset_bytes = off
            table[i].ndim = a.ndim
            for d in range(a.ndim):
                table[i].shape[d] = a.shape[d]
            chunks.append(a.reshape(-1))
            self.offsets.append(off // 4)
            off += a.size * 4
        blob = np.concatenate(chunks)
        _lib.check(self.h, self.L.a2m_train_init(self.h, blob.ctypes.data, blob.nbytes, table, n), "a2m_train_init")
        self.eng.weights_token = None   # the engine's inference weights now belong to the trainer
        self.eng.owner = self
        self.n_params = int(self.L.a2m_param_count(self.h))
        self.tdev = torch.device(f"cuda:{self.device}")
        self.grads = torch.zeros(self.n_params, dtype=torch.float32, device=self.tdev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.tdev)
        self.stats = torch.zeros(2, dtype=torch.float32, device=self.tdev)
        self._rope = None
        self.step_count = 0

    # ---- helpers
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.tdev).cuda_stream)

    def _rope_tensors(self, rope_freqs: RopeFreqs):
        if self._rope is None or self._rope[2] is not rope_freqs:
            t = self.torch
            cos = t.as_tensor(np.ascontiguousarray(rope_freqs.cos_freq, np.float32)).to(self.tdev)
            sin = t.as_tensor(np.ascontiguousarray(rope_freqs.sin_freq, np.float32)).to(self.tdev)
            self._rope = (cos, sin, rope_freqs)
        return self._rope[0], self._rope[1]

    def set_lr_multipliers(self, per_leaf: Optional[np.ndarray]):
        if per_leaf is None:
            _lib.check(self.h, self.L.a2m_set_lr_multipliers(self.h, None, 0), "a2m_set_lr_multipliers")
            return
        a = np.ascontiguousarray(per_leaf, np.float32)
        _lib.check(self.h, self.L.a2m_set_lr_multipliers(self.h, a.ctypes.data, a.size), "a2m_set_lr_multipliers")

    def set_dropout(self, rate: float, seed: int = 0):
        """transformer_dropout_rate of the following forward/backward pairs (model.py:30; the reference trains with 0.1)
        and the seed that stands in for the PRNG key of train.py:53."""
        _lib.check(self.h, self.L.a2m_set_dropout(self.h, float(rate), int(seed) & 0xFFFFFFFFFFFFFFFF), "a2m_set_dropout")

    # ---- compute_loss (train.py:48-62): forward with tape + backward, accumulating into self.grads / self.loss
    def zero_grad(self):
        self.grads.zero_()
        self.loss.zero_()

    def forward_backward(self, audio, labels, rope_freqs: RopeFreqs, scale: float = 1.0, want_logits: bool = False):
        t = self.torch
        if not (audio.is_cuda and labels.is_cuda):
            raise _lib.A2mError("training inputs must be CUDA tensors (no CPU path)")
        audio = audio.to(t.float32).contiguous()
        labels = labels.to(t.float32).contiguous()
        B = audio.shape[0]
        if tuple(audio.shape[1:]) != (2, 80000) or tuple(labels.shape) != (B, 250, 90):
            raise ValueError(f"audio (B, 2, 80000) / labels (B, 250, 90) expected, got {tuple(audio.shape)} / {tuple(labels.shape)}")
        cos, sin = self._rope_tensors(rope_freqs)
        logits = t.empty((B, 250, 90), dtype=t.float32, device=self.tdev) if want_logits else None
        rc = self.L.a2m_forward_train(self.h, audio.data_ptr(), B, cos.data_ptr(), sin.data_ptr(), cos.shape[0],
                                      logits.data_ptr() if want_logits else None, None, self._stream())
        _lib.check(self.h, rc, "a2m_forward_train")
        rc = self.L.a2m_backward(self.h, labels.data_ptr(), float(scale), self.grads.data_ptr(), self.loss.data_ptr(), self._stream())
        _lib.check(self.h, rc, "a2m_backward")
        self._keep = (audio, labels)   # the stem backward reads the audio asynchronously
        return logits

    def grad_buckets(self):
        """[(lo, hi)] element ranges of the gradient blob in the order they become final during a2m_backward: bucket 0 =
        final norm + transformer + decoder (ready once the transformer backward has run), bucket 1 = the CNN (ready at the end)."""
        out = []
        for k in range(int(self.L.a2m_grad_bucket_count(self.h))):
            lo, hi = C.c_size_t(), C.c_size_t()
            _lib.check(self.h, self.L.a2m_grad_bucket_range(self.h, k, C.byref(lo), C.byref(hi)), "a2m_grad_bucket_range")
            out.append((int(lo.value), int(hi.value)))
        return out

    def allreduce_grads(self, overlap: bool = True):
        """Data-parallel gradient exchange (train.py:238-244 shards the batch over devices): NCCL all-reduce (sum) over
        NVLink, then the mean over ranks; no-op without an initialised process group.  Bucketed and overlapped with the
        backward: call it right after the last forward_backward of the step -- bucket 0 (79 % of the bytes) is
        reduced on a communication stream as soon as a2m_backward's in-plan scatter has produced it, while the CNN backward
        still runs on the compute stream; the CNN bucket and the loss follow on the compute stream."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        torch = self.torch
        w = dist.get_world_size()
        buckets = self.grad_buckets() if overlap else []
        if len(buckets) == 2 and buckets[0][1] > buckets[0][0] and buckets[1][1] > buckets[1][0]:
            main = torch.cuda.current_stream(self.tdev)
            if getattr(self, "_comm", None) is None:
                self._comm = torch.cuda.Stream(self.tdev)
            comm = self._comm
            _lib.check(self.h, self.L.a2m_stream_wait_grad_bucket(self.h, 0, C.c_void_p(comm.cuda_stream)), "a2m_stream_wait_grad_bucket")
            tail = self.grads[buckets[0][0]:buckets[0][1]]
            with torch.cuda.stream(comm):
                dist.all_reduce(tail)
                tail.mul_(1.0 / w)
            head = self.grads[buckets[1][0]:buckets[1][1]]
            dist.all_reduce(head)
            head.mul_(1.0 / w)
            dist.all_reduce(self.loss)
            self.loss.mul_(1.0 / w)
            main.wait_stream(comm)
        else:
            allreduce_mean_([self.grads, self.loss])

    def optimizer_step(self, lr: float, cfg: OptimizerConfig, grad_divisor: float = 1.0):
        self.step_count += 1
        rc = self.L.a2m_adamw_step(self.h, self.grads.data_ptr(), float(lr), cfg.b1, cfg.b2, cfg.eps, cfg.weight_decay,
                                   float(grad_divisor), cfg.clip_norm, self.step_count, self.stats.data_ptr(), self._stream())
        _lib.check(self.h, rc, "a2m_adamw_step")

    # ---- compute_training_step (train.py:259-332)
    def training_step(self, audio, labels, rope_freqs: RopeFreqs, cfg: OptimizerConfig, lr: float, grad_scale: float = 1.0,
                      minibatch_size: Optional[int] = None, dropout_rate: float = 0.0, key: int = 0):
        """Minibatch scan with fp32 gradient accumulation, unscale by grad_scale x steps, all-reduce, AdamW + clip.
        Returns (loss, grads_valid, scaled_loss) as device tensors / lazily evaluated values (no host sync here)."""
        B = audio.shape[0]
        mb = B if minibatch_size is None else minibatch_size
        if B % mb != 0:
            raise ValueError("batch must be a multiple of the minibatch size")
        steps = B // mb
        self.zero_grad()
        for i in range(steps):
            self.set_dropout(dropout_rate, (int(key) * 0x9E3779B97F4A7C15 + self.step_count * 1315423911 + i) & 0xFFFFFFFFFFFFFFFF)
            self.forward_backward(audio[i * mb:(i + 1) * mb], labels[i * mb:(i + 1) * mb], rope_freqs, scale=grad_scale)
        self.allreduce_grads()
        self.optimizer_step(lr, cfg, grad_divisor=grad_scale * steps)
        scaled_loss = self.loss / steps
        return scaled_loss / grad_scale, self.stats[1] == 0, scaled_loss

    def train_pipelined(self, batches, rope_freqs: RopeFreqs, cfg: OptimizerConfig, lr_fn: Callable[[int], float], first_step: int = 1,
                        dropout_rate: float = 0.0, key: int = 0, grad_scale: float = 1.0):
        """Host-fed training loop (the reference's loop over a prefetching loader, train.py:340-380): `batches` is a sequence
        of (audio, labels) page-locked host tensors.  The H2D copy of batch i+1 runs on a copy stream while step i computes
        (two device buffers), and every step's loss is read back through a page-locked buffer that the host consumes one
        step later, so the host never waits on the step it has just enqueued.  Returns the list of per-step losses."""
        torch = self.torch
        main = torch.cuda.current_stream(self.tdev)
        n = len(batches)
        if n == 0:
            return []
        x0, y0 = batches[0]
        shapes = (tuple(x0.shape), tuple(y0.shape))
        pipe = getattr(self, "_pipe", None)
        if pipe is None or pipe["shapes"] != shapes:     # copy stream, device buffers, events and the page-locked loss slot live with the engine
            pipe = {"shapes": shapes, "copy": torch.cuda.Stream(self.tdev),
                    "bufs": [(torch.empty(shapes[0], dtype=torch.float32, device=self.tdev),
                              torch.empty(shapes[1], dtype=torch.float32, device=self.tdev)) for _ in range(2)],
                    "events": [[torch.cuda.Event(), torch.cuda.Event()] for _ in range(3)],
                    "pin_loss": torch.empty(2, dtype=torch.float32).pin_memory()}
            self._pipe = pipe
        copy, bufs, pin_loss = pipe["copy"], pipe["bufs"], pipe["pin_loss"]
        ready, consumed, loss_ev = pipe["events"]
        copy.wait_stream(main)                            # buffers may still be read by steps of an earlier call

        def prefetch(i):
            b = i & 1
            with torch.cuda.stream(copy):
                if i >= 2:
                    copy.wait_event(consumed[b])          # step i - 2 has finished reading this buffer
                bufs[b][0].copy_(batches[i][0], non_blocking=True)
                bufs[b][1].copy_(batches[i][1], non_blocking=True)
                ready[b].record(copy)

        losses = []
        prefetch(0)
        for i in range(n):
            b = i & 1
            if i + 1 < n:
                prefetch(i + 1)
            main.wait_event(ready[b])
            loss, _valid, _ = self.training_step(bufs[b][0], bufs[b][1], rope_freqs, cfg, lr_fn(first_step + i), grad_scale=grad_scale,
                                                 dropout_rate=dropout_rate, key=key)
            consumed[b].record(main)
            pin_loss[b:b + 1].copy_(loss.reshape(1), non_blocking=True)
            loss_ev[b].record(main)
            if i >= 1:
                loss_ev[b ^ 1].synchronize()
                losses.append(float(pin_loss[b ^ 1]))
        loss_ev[(n - 1) & 1].synchronize()
        losses.append(float(pin_loss[(n - 1) & 1]))
        return losses

    # ---- parameter access
    def params_flat(self):
        out = self.torch.empty(self.n_params, dtype=self.torch.float32, device=self.tdev)
        _lib.check(self.h, self.L.a2m_get_params(self.h, out.data_ptr(), self._stream()), "a2m_get_params")
        return out

    def _tree(self, flat) -> Dict[str, np.ndarray]:
        a = flat.detach().cpu().numpy()
        return {p: a[o:o + int(np.prod(s, dtype=np.int64))].reshape(s) for p, o, s in zip(self.paths, self.offsets, self.shapes)}

    def params_tree(self) -> Dict[str, np.ndarray]:
        return self._tree(self.params_flat())

    def grads_tree(self) -> Dict[str, np.ndarray]:
        return self._tree(self.grads)

    def profile_steps(self, which: int, repeats: int = 3):
        """Per-launch CUDA-event timings of the forward-with-tape (0) or backward (1) plan: (kernel, ms, flops, bytes)."""
        n = self.L.a2m_profile_train_steps(self.h, which, repeats, 0, None)
        if n < 0:
            _lib.check(self.h, n, "a2m_profile_train_steps")
        buf = (_lib.StepProfile * n)()
        m = self.L.a2m_profile_train_steps(self.h, which, repeats, n, buf)
        if m < 0:
            _lib.check(self.h, m, "a2m_profile_train_steps")
        return [(buf[i].kernel.decode(), float(buf[i].ms), float(buf[i].flops), float(buf[i].bytes)) for i in range(m)]

    def launch_count(self) -> int:
        return int(self.L.a2m_train_launch_count(self.h))


def compute_loss(model: OutputSequenceGenerator, state, audio, rope_freqs: RopeFreqs, expected_outputs, scale, key=None,
                 engine: Optional[TrainEngine] = None):
    """Reference call shape of compute_loss (train.py:48-62, under eqx.filter_value_and_grad(has_aux=True)):
    returns ((loss, state), grads) with grads keyed by pytree path.  `key` (an int) seeds dropout at the model's
    transformer_dropout_rate, as enable_dropout=True does in the reference; key=None runs without dropout."""
    eng = engine or TrainEngine(model)
    eng.zero_grad()
    eng.set_dropout(model_config["transformer_dropout_rate"] if key is not None else 0.0, 0 if key is None else int(key))
    eng.forward_backward(audio, expected_outputs, rope_freqs, scale=float(scale))
    return (eng.loss.clone(), state), eng.grads_tree()
