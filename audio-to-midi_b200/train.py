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
    """optax.join_schedules([linear 0 -> base over warmup, cosine_decay(base, steps)], [warmup])  (train.py:454-466).
    `step` is optax's update count, which starts at 0: schedule(0) = 0 is the learning rate of the FIRST update."""
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


def _dist_rank_world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


class TrainEngine:
    """Master parameters, AdamW state and activation tape on one GPU (a2m_train_init ...).

    Owns its OWN C handle (never the model's inference handle).  While it is alive, `model(...)` / `model.predict(...)` on the
    same device run on this handle, i.e. on the weights being trained (validation inside the loop, train.py:396-437);
    `sync_to_model()` copies the trained parameters back into the model's leaves (for save_checkpoint), `close()` ends the
    session."""

    def __init__(self, model: OutputSequenceGenerator, device: Optional[int] = None):
        import weakref
        import torch
        self.torch = torch
        self.device = _default_device() if device is None else device
        self.model = model
        self.closed = False
        self.eng = _Engine(self.device)
        self.L, self.h = self.eng.L, self.eng.h
        leaves = model.tree_leaves_with_path()
        self.paths = [p for p, _ in leaves]
        self.shapes = [tuple(np.shape(a)) for _, a in leaves]
        blob, table, self.offsets = _Engine.blob_and_table(leaves)
        _lib.check(self.h, self.L.a2m_train_init(self.h, blob.ctypes.data, blob.nbytes, table, len(table)), "a2m_train_init")
        self.model_version = model._version
        model._trainers[self.device] = weakref.ref(self)
        self.n_params = int(self.L.a2m_param_count(self.h))
        self.tdev = torch.device(f"cuda:{self.device}")
        self.grads = torch.zeros(self.n_params, dtype=torch.float32, device=self.tdev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.tdev)
        self.stats = torch.zeros(2, dtype=torch.float32, device=self.tdev)
        self._rope = None
        self.step_count = 0          # optimizer updates applied so far (optax's `count`)
        self.comm_ready = False

    def close(self):
        """Ends the training session and frees its handle; the model goes back to its own inference handle (with whatever
        leaves it holds: call sync_to_model() first to keep the trained weights)."""
        if not self.closed:
            self.closed = True
            self.eng.close()

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

    def forward_train(self, audio, rope_freqs: RopeFreqs, want_logits: bool = True, want_probs: bool = False):
        """a2m_forward_train: the dropout-enabled forward that records the tape (train.py:56-58).  Returns (logits, probs)
        device tensors (None where not wanted).  The audio tensor is kept alive until the backward has consumed it."""
        t = self.torch
        if not audio.is_cuda:
            raise _lib.A2mError("training inputs must be CUDA tensors (no CPU path)")
        audio = audio.to(t.float32).contiguous()
        B = audio.shape[0]
        if tuple(audio.shape[1:]) != (2, 80000):
            raise ValueError(f"audio (B, 2, 80000) expected, got {tuple(audio.shape)}")
        cos, sin = self._rope_tensors(rope_freqs)
        logits = t.empty((B, 250, 90), dtype=t.float32, device=self.tdev) if want_logits else None
        probs = t.empty((B, 250, 90), dtype=t.float32, device=self.tdev) if want_probs else None
        rc = self.L.a2m_forward_train(self.h, audio.data_ptr(), B, cos.data_ptr(), sin.data_ptr(), cos.shape[0],
                                      logits.data_ptr() if want_logits else None, probs.data_ptr() if want_probs else None, self._stream())
        _lib.check(self.h, rc, "a2m_forward_train")
        self._keep = (audio, None)   # the stem backward reads the audio asynchronously
        return logits, probs

    def backward(self, labels, scale: float = 1.0):
        """a2m_backward on the tape of the last forward_train: self.grads += d(mean_b sum BCE * scale)/dparams, self.loss += value."""
        t = self.torch
        labels = labels.to(t.float32).contiguous()
        B = self._keep[0].shape[0]
        if tuple(labels.shape) != (B, 250, 90) or not labels.is_cuda:
            raise ValueError(f"labels ({B}, 250, 90) on the GPU expected, got {tuple(labels.shape)}")
        rc = self.L.a2m_backward(self.h, labels.data_ptr(), float(scale), self.grads.data_ptr(), self.loss.data_ptr(), self._stream())
        _lib.check(self.h, rc, "a2m_backward")
        self._keep = (self._keep[0], labels)

    def backward_dlogits(self, dlogits):
        """a2m_backward_dlogits: self.grads += J^T dlogits for an arbitrary cotangent (the custom_vjp backward of a JAX host)."""
        t = self.torch
        dlogits = dlogits.to(t.float32).contiguous()
        rc = self.L.a2m_backward_dlogits(self.h, dlogits.data_ptr(), self.grads.data_ptr(), self._stream())
        _lib.check(self.h, rc, "a2m_backward_dlogits")
        self._keep = (self._keep[0], dlogits)

    def forward_backward(self, audio, labels, rope_freqs: RopeFreqs, scale: float = 1.0, want_logits: bool = False):
        if tuple(labels.shape) != (audio.shape[0], 250, 90):
            raise ValueError(f"audio (B, 2, 80000) / labels (B, 250, 90) expected, got {tuple(audio.shape)} / {tuple(labels.shape)}")
        logits, _ = self.forward_train(audio, rope_freqs, want_logits=want_logits)
        self.backward(labels, scale)
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

    def init_comm(self):
        """Creates the handle's own NCCL communicator over the ranks of the torch.distributed process group (rank 0's
        ncclGetUniqueId is broadcast through the group: the only thing torch.distributed does here is carry 128 bytes).
        Afterwards allreduce_grads() is ONE C call, a2m_allreduce_grads, with the bucket overlap inside the library."""
        import torch.distributed as dist
        rank, world = _dist_rank_world()
        if world == 1:
            return False
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            _lib.check(self.h, self.L.a2m_comm_unique_id(ident), "a2m_comm_unique_id")
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0)
        buf = (C.c_uint8 * 128).from_buffer_copy(box[0])
        _lib.check(self.h, self.L.a2m_comm_init(self.h, buf, world, rank), "a2m_comm_init")
        self.comm_ready = True
        return True

    def allreduce_grads(self, overlap: bool = True, backend: Optional[str] = None):
        """Data-parallel gradient exchange (train.py:238-244 shards the batch over devices): mean over ranks of self.grads and
        self.loss; no-op for one rank.  Call it right after the last forward_backward of the step.

        backend "a2m" (default once init_comm() has run): a2m_allreduce_grads -- ncclAllReduce per bucket inside the library,
        bucket 0 (79 % of the bytes) on the handle's communication stream under the CNN backward.  backend "torch": the same
        schedule with torch.distributed collectives (also what runs on CPU tensors / gloo in the tests)."""
        rank, world = _dist_rank_world()
        if world == 1:
            return
        backend = backend or ("a2m" if self.comm_ready else "torch")
        if backend == "a2m":
            _lib.check(self.h, self.L.a2m_allreduce_grads(self.h, None, self._stream()), "a2m_allreduce_grads")
            return
        import torch.distributed as dist
        torch = self.torch
        w = world
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
        """optax.adamw + clip_by_global_norm on the updates (train.py:324-325).  `lr` is the schedule at optax's PRE-increment
        count (0 for the first update: optax.scale_by_schedule evaluates schedule(count) before incrementing, so the first
        update of a warm-up schedule has lr 0); the bias correction uses count + 1.  On non-finite gradients the device step is
        a no-op (self.stats[1] != 0) -- the caller that notices may take the count back with `step_count -= 1`."""
        self.step_count += 1
        rc = self.L.a2m_adamw_step(self.h, self.grads.data_ptr(), float(lr), cfg.b1, cfg.b2, cfg.eps, cfg.weight_decay,
                                   float(grad_divisor), cfg.clip_norm, self.step_count, self.stats.data_ptr(), self._stream())
        _lib.check(self.h, rc, "a2m_adamw_step")

    # ---- compute_training_step (train.py:259-332)
    def training_step(self, audio, labels, rope_freqs: RopeFreqs, cfg: OptimizerConfig, lr: float, grad_scale: float = 1.0,
                      minibatch_size: Optional[int] = None, dropout_rate: float = 0.0, key: int = 0):
        """Minibatch scan with fp32 gradient accumulation, unscale by grad_scale x steps, all-reduce, AdamW + clip.
        Returns (loss, grads_valid, scaled_loss) as device tensors / lazily evaluated values (no host sync here).
        The dropout seed folds in the step, the minibatch index and the data-parallel RANK: the reference splits one key per
        sample of the GLOBAL batch (train.py:52-53), so ranks must not draw identical masks for their local samples."""
        B = audio.shape[0]
        mb = B if minibatch_size is None else minibatch_size
        if B % mb != 0:
            raise ValueError("batch must be a multiple of the minibatch size")
        steps = B // mb
        rank, _world = _dist_rank_world()
        self.zero_grad()
        for i in range(steps):
            seed = (int(key) * 0x9E3779B97F4A7C15 + self.step_count * 1315423911 + i + (rank + 1) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
            self.set_dropout(dropout_rate, seed)
            self.forward_backward(audio[i * mb:(i + 1) * mb], labels[i * mb:(i + 1) * mb], rope_freqs, scale=grad_scale)
        self.allreduce_grads()
        self.optimizer_step(lr, cfg, grad_divisor=grad_scale * steps)
        scaled_loss = self.loss / steps
        return scaled_loss / grad_scale, self.stats[1] == 0, scaled_loss

    def train_pipelined(self, batches, rope_freqs: RopeFreqs, cfg: OptimizerConfig, lr_fn: Callable[[int], float], first_step: int = 0,
                        dropout_rate: float = 0.0, key: int = 0, grad_scale: float = 1.0, return_valid: bool = False):
        """Host-fed training loop (the reference's loop over a prefetching loader, train.py:340-380): `batches` is a sequence
        of (audio, labels) page-locked host tensors.  The H2D copy of batch i+1 runs on a copy stream while step i computes
        (two device buffers), and every step's loss and grads_valid flag are read back through a page-locked buffer that the
        host consumes one step later, so the host never waits on the step it has just enqueued.

        lr_fn is called with optax's 0-based count: lr_fn(first_step + i) for the i-th step of this call.  When a step reports
        non-finite gradients (the device update was a no-op) the loop does what train.py:369-377 does without needing the
        snapshot: the loss scale `grad_scale` is halved for the following steps and the update count is taken back.
        Returns the per-step losses (and, with return_valid, the per-step validity flags and the final grad_scale)."""
        torch = self.torch
        main = torch.cuda.current_stream(self.tdev)
        n = len(batches)
        if n == 0:
            return ([], [], grad_scale) if return_valid else []
        x0, y0 = batches[0]
        shapes = (tuple(x0.shape), tuple(y0.shape))
        pipe = getattr(self, "_pipe", None)
        if pipe is None or pipe["shapes"] != shapes:     # copy stream, device buffers, events and the page-locked loss slot live with the engine
            pipe = {"shapes": shapes, "copy": torch.cuda.Stream(self.tdev),
                    "bufs": [(torch.empty(shapes[0], dtype=torch.float32, device=self.tdev),
                              torch.empty(shapes[1], dtype=torch.float32, device=self.tdev)) for _ in range(2)],
                    "events": [[torch.cuda.Event(), torch.cuda.Event()] for _ in range(3)],
                    "pin_loss": torch.empty(4, dtype=torch.float32).pin_memory()}
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

        losses, valids = [], []
        skipped = 0

        def consume(b):
            nonlocal grad_scale, skipped
            loss_ev[b].synchronize()
            losses.append(float(pin_loss[2 * b]))
            ok = float(pin_loss[2 * b + 1]) == 0.0
            valids.append(ok)
            if not ok:                                    # train.py:369-377: halve the loss scale; the update was a device no-op
                grad_scale = grad_scale / 2.0
                self.step_count -= 1
                skipped += 1

        prefetch(0)
        for i in range(n):
            b = i & 1
            if i + 1 < n:
                prefetch(i + 1)
            main.wait_event(ready[b])
            loss, _valid, _ = self.training_step(bufs[b][0], bufs[b][1], rope_freqs, cfg, lr_fn(first_step + i - skipped), grad_scale=grad_scale,
                                                 dropout_rate=dropout_rate, key=key)
            consumed[b].record(main)
            pin_loss[2 * b:2 * b + 1].copy_(loss.reshape(1), non_blocking=True)
            pin_loss[2 * b + 1:2 * b + 2].copy_(self.stats[1:2], non_blocking=True)
            loss_ev[b].record(main)
            if i >= 1:
                consume(b ^ 1)
        consume((n - 1) & 1)
        return (losses, valids, grad_scale) if return_valid else losses

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

    def set_params_flat(self, flat):
        """Overwrites the master parameters (blob layout, device tensor) and re-packs the kernel images."""
        flat = flat.to(self.tdev, self.torch.float32).contiguous()
        if flat.numel() != self.n_params:
            raise ValueError("parameter blob has the wrong size")
        _lib.check(self.h, self.L.a2m_set_params(self.h, flat.data_ptr(), self._stream()), "a2m_set_params")
        self._keep_params = flat

    def snapshot(self):
        """(params, mu, nu, step_count) device copies: the in-memory snapshot of train.py:348-353 (copy_pytree every 100 steps)."""
        t = self.torch
        m = t.empty(self.n_params, dtype=t.float32, device=self.tdev)
        v = t.empty_like(m)
        _lib.check(self.h, self.L.a2m_get_opt_state(self.h, m.data_ptr(), v.data_ptr(), self._stream()), "a2m_get_opt_state")
        return self.params_flat(), m, v, self.step_count

    def restore(self, snap):
        """Rolls the session back to a snapshot() (train.py:369-377, the recovery from non-finite gradients)."""
        params, m, v, step = snap
        self.set_params_flat(params)
        _lib.check(self.h, self.L.a2m_set_opt_state(self.h, m.data_ptr(), v.data_ptr(), self._stream()), "a2m_set_opt_state")
        self.step_count = int(step)

    def sync_to_model(self) -> OutputSequenceGenerator:
        """Copies the trained master parameters into the model's leaves (what `model = eqx.apply_updates(model, updates)`
        hands back in the reference, train.py:325), so that infer.save_checkpoint(model, ...) stores the trained weights.
        Inference through the model keeps running on this session's handle."""
        self.model.load_leaves(self.params_tree())
        self.model_version = self.model._version
        return self.model

    def save_checkpoint(self, directory: str, step: Optional[int] = None):
        """train.py:384-394: saves the CURRENT (trained) parameters."""
        from .infer import save_checkpoint
        return save_checkpoint(self.sync_to_model(), directory, self.step_count if step is None else step)

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
    eng = engine or model._live_trainer(_default_device()) or TrainEngine(model)
    eng.zero_grad()
    eng.set_dropout(model_config["transformer_dropout_rate"] if key is not None else 0.0, 0 if key is None else int(key))
    eng.forward_backward(audio, expected_outputs, rope_freqs, scale=float(scale))
    return (eng.loss.clone(), state), eng.grads_tree()
