"""Host-side mirror of the reference model interface (reference model.py), backed by the CUDA library.

Same class names, field names, field ORDER and leaf shapes as the reference's equinox modules, so the
flattened pytree (``tree_leaves_with_path``) lines up leaf-for-leaf with the reference's checkpoint
pytree (model.py:673-678 and the classes below it).  Same call signatures:

    model = OutputSequenceGenerator(model_config, key)
    (logits, probs), state = model(samples, state, rope_freqs, key=None, enable_dropout=False)   # model.py:740-769
    logits, probs = model.predict(state, samples, rope_freqs)                                     # model.py:771-773

The reference model is unbatched and callers ``jax.vmap`` it (infer.py:40).  JAX does not exist in this
image, so the batch axis is native here: ``samples`` may be (2, N) or (B, 2, N); ``vmap(model.predict,
in_axes=(None, 0, None))`` is provided as a thin adapter that forwards the batched array unchanged.

Arrays: numpy in -> numpy out (host path, copies inside the C call); torch CUDA tensors in -> torch CUDA
tensors out (device path, enqueued on the current torch stream, no host sync).  No CPU compute path
exists: without the CUDA library / an sm_100 GPU every call raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional

import numpy as np

from . import _lib
from .rope import RopeFreqs

HOST_SLOTS = 4                    # include/a2m.h A2M_HOST_SLOTS: batches in flight on the pipelined host path
MIDI_EVENT_VOCCAB_SIZE = 90       # audio_to_midi_dataset.py:26
MODEL_AUDIO_LENGTH = 5.0          # audio_to_midi_dataset.py:28
SAMPLE_RATE = 16000               # audio_to_midi_dataset.py:111

model_config = {                   # model.py:20-34
    "dims": [4 * (2 ** i) for i in range(7)],
    "depths": [3, 3, 3, 3, 3, 21, 3],
    "cnn_hidden_expansion": 2.0,
    "num_transformer_layers": 8,
    "num_transformer_heads": 4,
    "attention_size": 64,
    "compressed_attention_q_size": 64,
    "compressed_attention_kv_size": 64,
    "transformer_dropout_rate": 0.1,
    "transformer_hidden_expansion": 2.0,
    "sdd_rate": 0.1,
}


def get_model_metadata():         # model.py:36-41
    return {"model": model_config,
            "data_prep": {"sample_rate": SAMPLE_RATE, "audio_length": MODEL_AUDIO_LENGTH}}


# Operand format of inference handles unless change_fp_precision says otherwise: "f16" (IEEE binary16 tensor-core operands, the
# libaudio2midi_b200_f16.so build; measured on B200: max |dprob| 2.1e-3 against the fp32 twin, 8x below bf16's 1.8e-2, at the same
# speed) or "bf16".  Accumulators, the residual stream and LN / softmax statistics are fp32 either way.  Training is bf16.
import os as _os
DEFAULT_INFERENCE_PRECISION = _os.environ.get("A2M_INFER_PRECISION", "f16")


def change_fp_precision(model, dtype):
    """infer.py:27-32: the reference casts every inexact leaf (fp32 for inference, infer.py:234; fp16 for training, train.py:36-37).
    Here the leaves stay fp32 masters and `dtype` picks the tensor-core OPERAND format the model's inference handles are built
    with: float16 -> "f16" (11-bit significand), bfloat16 -> "bf16" (8-bit); float32 maps to "f16", the closest the sm_100a
    kind::f16 path offers (its 2^-11 operand rounding is TF32's).  Returns the model (changed in place)."""
    name = getattr(dtype, "__name__", None) or getattr(dtype, "name", None) or str(dtype)
    name = name.replace("torch.", "").replace("jnp.", "")
    table = {"float16": "f16", "f16": "f16", "half": "f16", "float32": "f16", "f32": "f16", "bfloat16": "bf16", "bf16": "bf16"}
    if name not in table:
        raise ValueError(f"unsupported precision {dtype!r}")
    model.precision = table[name]
    return model


# ----------------------------------------------------------------------------------------------- pytree
class Module:
    """Minimal stand-in for eqx.Module: ordered fields, leaves are numpy arrays (or None)."""
    _fields: tuple = ()

    def tree_leaves_with_path(self, prefix=""):
        out = []
        for name in self._fields:
            out.extend(_flatten(getattr(self, name), f"{prefix}{name}"))
        return out


def _flatten(v, path):
    if v is None:
        return []
    if isinstance(v, Module):
        return v.tree_leaves_with_path(path + ".")
    if isinstance(v, (list, tuple)):
        out = []
        for i, x in enumerate(v):
            out.extend(_flatten(x, f"{path}.{i}"))
        return out
    if isinstance(v, (np.ndarray, np.floating)):
        return [(path, v)]
    return []  # static fields (ints, bools)


def _set_by_path(obj, path, value):
    parts = path.split(".")
    for p in parts[:-1]:
        obj = obj[int(p)] if p.isdigit() else getattr(obj, p)
    last = parts[-1]
    cur = obj[int(last)] if last.isdigit() else getattr(obj, last)
    value = np.asarray(value, dtype=np.float32)
    if np.shape(cur) != value.shape:
        raise ValueError(f"leaf {path}: shape {value.shape} does not match {np.shape(cur)}")
    if last.isdigit():
        obj[int(last)] = value
    else:
        setattr(obj, last, value)


class _Rng:
    def __init__(self, key):
        seed = 0 if key is None else int(np.asarray(key).ravel()[-1])
        self.g = np.random.Generator(np.random.PCG64(seed))

    def uniform(self, shape, fan_in):
        lim = 1.0 / np.sqrt(fan_in)
        return self.g.uniform(-lim, lim, size=shape).astype(np.float32)


class Conv1d(Module):              # eqx.nn.Conv1d: weight (out, in/groups, k), bias (out, 1)
    _fields = ("weight", "bias")

    def __init__(self, rng, cin, cout, k, groups=1, lead=()):
        fan_in = (cin // groups) * k
        self.weight = rng.uniform(lead + (cout, cin // groups, k), fan_in)
        self.bias = rng.uniform(lead + (cout, 1), fan_in)


class Linear(Module):              # eqx.nn.Linear: weight (out, in), bias (out,)
    _fields = ("weight", "bias")

    def __init__(self, rng, cin, cout, use_bias=True, lead=()):
        self.weight = rng.uniform(lead + (cout, cin), cin)
        self.bias = rng.uniform(lead + (cout,), cin) if use_bias else None


class LayerNorm(Module):           # eqx.nn.LayerNorm: weight, bias
    _fields = ("weight", "bias")

    def __init__(self, n, lead=()):
        self.weight = np.ones(lead + (n,), np.float32)
        self.bias = np.zeros(lead + (n,), np.float32)


class StochasticDepthDropout(Module):   # model.py:49-81; p is an array leaf in the reference (model.py:694,710)
    _fields = ("p",)

    def __init__(self, p):
        self.p = np.float32(p)
        self.inference = False


class Stem(Module):                # model.py:84-100
    _fields = ("conv", "norm")

    def __init__(self, rng, channels, kernel_size=5):
        self.conv = Conv1d(rng, 2, channels, kernel_size)
        self.norm = LayerNorm(channels)


class Downsample(Module):          # model.py:102-118
    _fields = ("conv", "norm")

    def __init__(self, rng, cin, cout):
        self.conv = Conv1d(rng, cin, cout, 2)
        self.norm = LayerNorm(cin)


class Block(Module):               # model.py:120-167
    _fields = ("depth_conv", "point_conv_1", "point_conv_2", "stochastic_depth_dropout", "norm", "gamma")

    def __init__(self, rng, channels, hidden_dim, sdd_rate, kernel_size=7):
        self.depth_conv = Conv1d(rng, channels, channels, kernel_size, groups=channels)
        self.norm = LayerNorm(channels)
        self.point_conv_1 = Conv1d(rng, channels, hidden_dim, 1)
        self.point_conv_2 = Conv1d(rng, hidden_dim, channels, 1)
        self.stochastic_depth_dropout = StochasticDepthDropout(sdd_rate)
        self.gamma = np.full((channels,), 1e-6, np.float32)   # layer scale, model.py:157-158


class Sequential(Module):          # eqx.nn.Sequential: field `layers`
    _fields = ("layers",)

    def __init__(self, layers):
        self.layers = list(layers)


class Decoder(Module):             # model.py:169-198
    _fields = ("decoder_pooling", "norm")

    def __init__(self, rng, dim):
        self.decoder_pooling = Linear(rng, dim, MIDI_EVENT_VOCCAB_SIZE)
        self.norm = LayerNorm(dim)


class FeedForwardBlock(Module):    # model.py:200-238 (dropout has no array leaves)
    _fields = ("attention_to_intermediate_proj", "intermediate_to_attention_proj")

    def __init__(self, rng, hidden, inter, lead):
        self.attention_to_intermediate_proj = Linear(rng, hidden, 2 * inter, lead=lead)
        self.intermediate_to_attention_proj = Linear(rng, inter, hidden, lead=lead)


class SelfAttention(Module):       # model.py:260-374; query_down_proj is None in the default config
    _fields = ("query_down_proj", "query_up_proj", "kv_down_proj", "key_up_proj", "value_up_proj", "output_proj")

    def __init__(self, rng, d, heads, hd, ckv, lead):
        self.query_down_proj = None
        self.query_up_proj = Linear(rng, d, heads * hd, use_bias=False, lead=lead)
        self.kv_down_proj = Linear(rng, d, ckv, use_bias=False, lead=lead)
        self.key_up_proj = Linear(rng, ckv, heads * hd, use_bias=False, lead=lead)
        self.value_up_proj = Linear(rng, ckv, heads * hd, use_bias=False, lead=lead)
        self.output_proj = Linear(rng, heads * hd, d, use_bias=False, lead=lead)
        self.num_heads = heads


class LocalSelfAttention(Module):  # model.py:377-471
    _fields = ("self_attention",)

    def __init__(self, rng, context_length, d, heads, hd, ckv, lead):
        self.context_length = context_length
        self.self_attention = SelfAttention(rng, d, heads, hd, ckv, lead)


class TransformerLayer(Module):    # model.py:474-556
    _fields = ("attention_norm", "attention_block", "feed_forward_norm", "feed_forward_block")

    def __init__(self, rng, d, heads, hd, ckv, inter, lead, context_window=None):
        if context_window is not None:
            self.attention_block = LocalSelfAttention(rng, context_window, d, heads, hd, ckv, lead)
        else:
            self.attention_block = SelfAttention(rng, d, heads, hd, ckv, lead)
        self.attention_norm = LayerNorm(d, lead)
        self.feed_forward_block = FeedForwardBlock(rng, d, inter, lead)
        self.feed_forward_norm = LayerNorm(d, lead)


class AlternatingLocalAndGlobalAttention(Module):   # model.py:559-612
    _fields = ("local_attention", "global_attention")

    def __init__(self, rng, d, heads, hd, ckv, inter, lead):
        self.local_attention = TransformerLayer(rng, d, heads, hd, ckv, inter, lead, context_window=16)
        self.global_attention = TransformerLayer(rng, d, heads, hd, ckv, inter, lead)


class TransformerStack(Module):    # model.py:615-670: `layers` is ONE module whose leaves carry a leading axis
    _fields = ("layers",)

    def __init__(self, rng, d, num_layers, heads, hd, ckv, inter):
        self.num_layers = num_layers
        self.layers = AlternatingLocalAndGlobalAttention(rng, d, heads, hd, ckv, inter, lead=(num_layers,))


# ----------------------------------------------------------------------------------------------- engine
def default_config_struct(device: int) -> "_lib.A2mConfig":
    """model_config (model.py:20-34) as the A2mConfig of include/a2m.h."""
    c = _lib.A2mConfig()
    c.device = device
    c.num_stages = len(model_config["dims"])
    for i, (d, n) in enumerate(zip(model_config["dims"], model_config["depths"])):
        c.dims[i], c.depths[i] = d, n
    c.cnn_hidden_expansion_x2 = int(round(model_config["cnn_hidden_expansion"] * 2))
    c.num_transformer_layers = model_config["num_transformer_layers"]
    c.num_transformer_heads = model_config["num_transformer_heads"]
    c.attention_size = model_config["attention_size"]
    c.compressed_attention_kv_size = model_config["compressed_attention_kv_size"]
    c.transformer_intermediate = int(model_config["dims"][-1] * model_config["transformer_hidden_expansion"])
    c.use_graph = c.use_pdl = -1
    return c


class _Engine:
    """One C handle (a2m_create_ex): the weights arena, workspace and launch plans of ONE model on ONE device.  A model owns
    its engines (`OutputSequenceGenerator._engines`), a TrainEngine owns its own: handles are never shared or stolen, so a
    model and a trainer -- or two models -- can be resident on the same GPU at the same time."""

    def __init__(self, device: int, precision: str = "bf16"):
        self.L = _lib.lib(precision)
        self.precision = precision
        h = C.c_void_p()
        cfg = default_config_struct(device)
        rc = self.L.a2m_create_ex(C.byref(cfg), C.byref(h))
        self.h = h
        if rc != 0:
            msg = self.L.a2m_last_error(h).decode() if h else ""
            if h:
                self.L.a2m_destroy(h)
                self.h = None
            raise _lib.A2mError(f"a2m_create_ex(device={device}) failed with code {rc} {msg}: an sm_100 (B200) GPU and "
                                "the CUDA library are required; there is no CPU fallback")
        self.device = device
        self.weights_token = None

    def close(self):
        if getattr(self, "h", None):
            self.L.a2m_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def blob_and_table(leaves):
        """[(path, array)] -> (contiguous fp32 blob, A2mLeafDesc table, element offsets)."""
        n = len(leaves)
        table = (_lib.LeafDesc * n)()
        chunks, off, offsets = [], 0, []
        for i, (path, arr) in enumerate(leaves):
            a = np.ascontiguousarray(arr, dtype=np.float32)
            table[i].path = path.encode()
            table[i].offset_bytes = off
            table[i].ndim = a.ndim
            for d in range(a.ndim):
                table[i].shape[d] = a.shape[d]
            chunks.append(a.reshape(-1))
            offsets.append(off // 4)
            off += a.size * 4
        return np.concatenate(chunks), table, offsets

    def load(self, model: "OutputSequenceGenerator"):
        blob, table, _ = self.blob_and_table(model.tree_leaves_with_path())
        _lib.check(self.h, self.L.a2m_load_weights(self.h, blob.ctypes.data, blob.nbytes, table, len(table)), "a2m_load_weights")
        self.weights_token = model._version


class OutputSequenceGenerator(Module):   # model.py:673-773
    _fields = ("layers", "norm", "transformer_projection", "transformer", "decoder")

    def __init__(self, conf: Dict[str, Any], key=None):
        rng = _Rng(key)
        dims, depths = conf["dims"], conf["depths"]
        if list(dims) != model_config["dims"] or list(depths) != model_config["depths"] or \
                conf.get("transformer_hidden_dim", dims[-1]) != dims[-1]:
            raise NotImplementedError("the CUDA kernels are specialised for the reference's default model_config")
        hidden = [int(d * conf["cnn_hidden_expansion"]) for d in dims]
        sdd = np.linspace(0.0, conf["sdd_rate"], sum(depths))
        self.layers, k = [], 0
        for i in range(len(dims)):
            first = Stem(rng, dims[0]) if i == 0 else Downsample(rng, dims[i - 1], dims[i])
            blocks = [Block(rng, dims[i], hidden[i], sdd[k + j]) for j in range(depths[i])]
            k += depths[i]
            self.layers.append(Sequential([first, *blocks]))
        self.norm = LayerNorm(dims[-1])
        self.transformer_projection = None
        d = dims[-1]
        self.transformer = TransformerStack(rng, d, conf["num_transformer_layers"], conf["num_transformer_heads"],
                                            conf["attention_size"], conf["compressed_attention_kv_size"],
                                            int(d * conf["transformer_hidden_expansion"]))
        self.decoder = Decoder(rng, d)
        self._version = 0
        self._rope_cache = {}
        self._out_ring = {}
        self._engines = {}        # (device, precision) -> _Engine owned by this model
        self.precision = DEFAULT_INFERENCE_PRECISION   # tensor-core operand format of this model's inference (change_fp_precision)
        self._trainers = {}       # device -> weakref to the live TrainEngine built from this model (train.py)
        self._own_trainers = {}   # device -> TrainEngine created implicitly by model(..., enable_dropout=True)
        self._lanes = {}          # (device, precision) -> streams + workspaces of predict_many

    # -- pytree helpers (what eqx.tree_at / tree_deserialise_leaves would be used for)
    def load_leaves(self, leaves: Dict[str, np.ndarray]):
        """Overwrite parameters from {dotted key path: array}; every leaf of the pytree must be present."""
        mine = [p for p, _ in self.tree_leaves_with_path()]
        missing = [p for p in mine if p not in leaves]
        if missing:
            raise KeyError(f"missing leaves: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        for p in mine:
            _set_by_path(self, p, leaves[p])
        self._version += 1
        return self

    def invalidate(self):
        """Call after mutating leaves in place so the next forward re-uploads the weights."""
        self._version += 1

    # -- forward
    def _live_trainer(self, device: int):
        ref = self._trainers.get(device)
        t = ref() if ref is not None else None
        if t is None or t.closed or t.model_version != self._version:
            return None       # no trainer, or the model's leaves were replaced after the trainer was built
        return t

    def _engine(self, device: int) -> _Engine:
        """The handle inference runs on.  While a TrainEngine built from this model is alive on `device`, that is the
        TRAINER's handle: a2m_forward reads the arena the optimizer re-packs after every step, so validation inside a training
        loop (train.py:396-437) sees the current weights and never disturbs the training session."""
        t = self._live_trainer(device)
        if t is not None:
            return t.eng
        eng = self._engines.get((device, self.precision))
        if eng is None:
            eng = self._engines[(device, self.precision)] = _Engine(device, self.precision)
        if eng.weights_token != self._version:
            eng.load(self)
        return eng

    def __call__(self, samples, state, rope_freqs: RopeFreqs, key=None, enable_dropout: bool = False):
        if enable_dropout:
            logits, probs = self._forward_train(samples, rope_freqs, key)
        else:
            logits, probs = self._forward(samples, rope_freqs)
        return (logits, probs), state

    def _forward_train(self, samples, rope_freqs, key):
        """Training-mode forward, the call shape of train.py:56-58 (`model(audio, state, rope_freqs, key, True)` under vmap):
        dropout at transformer_dropout_rate on the attention weights and the FFN output, masks seeded by `key` (an int, or
        an array of per-sample keys that is folded into one seed: the masks are a counter-based hash of (seed, site, element),
        not jax's threefry).  Runs a2m_forward_train on this model's TrainEngine, so a following
        `TrainEngine.backward(labels)` differentiates exactly this forward."""
        from .train import TrainEngine
        if not type(samples).__module__.startswith("torch") or not samples.is_cuda:
            raise _lib.A2mError("the training-mode forward takes CUDA tensors (no CPU path)")
        import torch
        single = samples.ndim == 2
        x = samples.reshape((-1, 2, 80000)).to(torch.float32).contiguous()
        dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
        t = self._live_trainer(dev)
        if t is None:
            t = self._own_trainers[dev] = TrainEngine(self, dev)      # kept alive by the model: the backward needs its tape
        seed = fold_key(key)
        t.set_dropout(model_config["transformer_dropout_rate"], seed)
        logits, probs = t.forward_train(x, rope_freqs, want_probs=True)
        return (logits[0], probs[0]) if single else (logits, probs)

    def predict(self, state, samples, rope_freqs: RopeFreqs):
        (logits, probs), _ = self(samples, state, rope_freqs, None)
        return logits, probs

    def _forward(self, samples, rope_freqs):
        is_torch = type(samples).__module__.startswith("torch")
        shape = tuple(samples.shape)
        single = len(shape) == 2
        if (len(shape) not in (2, 3)) or shape[-2:] != (2, 80000):
            raise ValueError(f"samples must be (2, 80000) or (B, 2, 80000), got {shape}")
        B = 1 if single else shape[0]
        if B == 0:      # an empty batch maps to empty outputs, as jax.vmap over a zero-length axis does; no launch, no device needed
            if is_torch:
                import torch
                z = torch.empty((0, 250, 90), dtype=torch.float32, device=samples.device)
                return z, z.clone()
            return np.empty((0, 250, 90), np.float32), np.empty((0, 250, 90), np.float32)
        if is_torch:
            import torch
            if not samples.is_cuda:
                raise _lib.A2mError("torch input must live on a CUDA device (no CPU path); pass numpy for the host API")
            dev = samples.device.index if samples.device.index is not None else torch.cuda.current_device()
            eng = self._engine(dev)
            x = samples.to(torch.float32).contiguous()
            hit = self._rope_cache.get(dev)
            if hit is None or hit[2] is not rope_freqs:     # identity of the live object, never a recycled id()
                cos = torch.as_tensor(np.ascontiguousarray(rope_freqs.cos_freq, np.float32)).to(samples.device)
                sin = torch.as_tensor(np.ascontiguousarray(rope_freqs.sin_freq, np.float32)).to(samples.device)
                hit = self._rope_cache[dev] = (cos, sin, rope_freqs)
            cos, sin, _ = hit
            logits = torch.empty((B, 250, 90), dtype=torch.float32, device=samples.device)
            probs = torch.empty_like(logits)
            stream = torch.cuda.current_stream(samples.device).cuda_stream
            rc = eng.L.a2m_forward(eng.h, x.data_ptr(), B, cos.data_ptr(), sin.data_ptr(), cos.shape[0],
                                   logits.data_ptr(), probs.data_ptr(), None, 0, C.c_void_p(stream))
            _lib.check(eng.h, rc, "a2m_forward", eng.L)
            return (logits[0], probs[0]) if single else (logits, probs)
        x = np.ascontiguousarray(samples, dtype=np.float32)
        eng = self._engine(_default_device())
        cos = np.ascontiguousarray(rope_freqs.cos_freq, np.float32)
        sin = np.ascontiguousarray(rope_freqs.sin_freq, np.float32)
        logits = np.empty((B, 250, 90), np.float32)
        probs = np.empty((B, 250, 90), np.float32)
        rc = eng.L.a2m_forward_host(eng.h, x.ctypes.data, B, cos.ctypes.data, sin.ctypes.data, cos.shape[0],
                                    logits.ctypes.data, probs.ctypes.data)
        _lib.check(eng.h, rc, "a2m_forward_host", eng.L)
        return (logits[0], probs[0]) if single else (logits, probs)

    def predict_many(self, state, batches, rope_freqs: RopeFreqs, lanes: int = 2):
        """Throughput form of `vmap(model.predict)` (infer.py:40) for SEVERAL device-resident batches -- the chunks of a long clip
        (infer.py:339), a validation set, consecutive serving batches.  `batches` is a sequence of torch CUDA tensors (B_i, 2, 80000);
        returns [(logits_i, probs_i)] in order.

        The batches alternate over `lanes` CUDA streams, each with its own workspace and launch plan (a2m_forward with a caller
        workspace), so that two consecutive, independent batches are in flight at once.  At 64 windows every kernel of the plan is at
        most one wave of CTAs and most of them are chains of dependent phases; a second batch fills the SMs and the wave tails
        the first leaves idle: 1.32 ms per 64-window batch with two lanes against 1.55 ms one after the other (B200, round 2).
        Stream semantics are those of one call: the work is ordered after everything already enqueued on the current stream, and
        the current stream waits for all of it before anything enqueued afterwards runs.  No host synchronisation."""
        import torch
        batches = list(batches)
        if len(batches) <= 1 or lanes <= 1:
            return [self.predict(state, x, rope_freqs) for x in batches]
        x0 = batches[0]
        if not (type(x0).__module__.startswith("torch") and x0.is_cuda):
            raise _lib.A2mError("predict_many takes torch CUDA tensors; host arrays go through predict_pipelined")
        dev = x0.device.index if x0.device.index is not None else torch.cuda.current_device()
        eng = self._engine(dev)
        main = torch.cuda.current_stream(x0.device)
        xs = []
        for x in batches:
            if x.ndim != 3 or tuple(x.shape[1:]) != (2, 80000) or x.device != x0.device:
                raise ValueError(f"batches must be CUDA tensors (B, 2, 80000) on one device, got {tuple(x.shape)}")
            xs.append(x.to(torch.float32).contiguous())
        hit = self._rope_cache.get(dev)
        if hit is None or hit[2] is not rope_freqs:
            cos = torch.as_tensor(np.ascontiguousarray(rope_freqs.cos_freq, np.float32)).to(x0.device)
            sin = torch.as_tensor(np.ascontiguousarray(rope_freqs.sin_freq, np.float32)).to(x0.device)
            hit = self._rope_cache[dev] = (cos, sin, rope_freqs)
        cos, sin, _ = hit
        max_b = max(1, max(int(x.shape[0]) for x in xs))
        key = (dev, self.precision)
        st = self._lanes.get(key)
        if st is None or st["eng"] is not eng or st["cap"] < max_b or len(st["streams"]) < lanes:
            need = int(eng.L.a2m_workspace_bytes(eng.h, max_b, 0))
            raw = [torch.zeros(need + 1024, dtype=torch.uint8, device=x0.device) for _ in range(lanes)]
            st = self._lanes[key] = {"eng": eng, "cap": max_b, "bytes": need, "raw": raw,
                                     "ws": [(t.data_ptr() + 1023) & ~1023 for t in raw],
                                     "streams": [torch.cuda.Stream(x0.device) for _ in range(lanes)]}
        # every output is allocated on the current stream BEFORE the fork, so no tensor ever crosses allocator streams; ONE
        # allocation for all batches (views are returned): a cudaMalloc per batch inside the loop would serialise the lanes
        total = sum(int(x.shape[0]) for x in xs)
        big = torch.empty((2, total, 250, 90), dtype=torch.float32, device=x0.device)
        outs, off = [], 0
        for x in xs:
            n = int(x.shape[0])
            outs.append((big[0, off:off + n], big[1, off:off + n]))
            off += n
        fork = torch.cuda.Event()
        fork.record(main)
        for s in st["streams"][:lanes]:
            s.wait_event(fork)
        lane = -1
        for x, (lg, pr) in zip(xs, outs):
            if int(x.shape[0]) == 0:      # an empty batch: empty views, no launch
                continue
            lane = (lane + 1) % lanes
            rc = eng.L.a2m_forward(eng.h, x.data_ptr(), int(x.shape[0]), cos.data_ptr(), sin.data_ptr(), cos.shape[0], lg.data_ptr(),
                                   pr.data_ptr(), C.c_void_p(st["ws"][lane]), st["bytes"], C.c_void_p(st["streams"][lane].cuda_stream))
            _lib.check(eng.h, rc, "a2m_forward", eng.L)
        for s in st["streams"][:lanes]:
            main.wait_stream(s)
        self._keep_many = xs          # the lanes read the (possibly converted) inputs asynchronously
        return outs

    def predict_pipelined(self, batches, rope_freqs: RopeFreqs, state=None, copy: bool = False, want_logits: bool = True,
                          probs_dtype=np.float32):
        """Generator over an iterable of host batches (B, 2, 80000): yields (logits, probs) per batch, in order,
        keeping four batches in flight -- two computing on the two compute lanes (their kernels overlap each other), one uploading,
        one downloading (a2m_submit_host_ex / a2m_collect_host).  Use ``pinned_empty`` arrays for the inputs to make the copies
        truly asynchronous.  Outputs live in a ring of five page-locked buffers (page-locking is expensive, so
        they are allocated once per batch size): a yielded pair stays valid until the next pair has been
        yielded; pass copy=True to get private copies instead.

        Bytes over PCIe: float16 batches are uploaded as they are (lossless for audio normalised by load_full_audio, which
        rounds to f16, python.rs:235-264) and widened on the device; want_logits=False skips the logits read-back (infer.py:41
        keeps only the probabilities; None is yielded in their place); probs_dtype=np.float16 halves the other half."""
        eng = self._engine(_default_device())
        cos = np.ascontiguousarray(rope_freqs.cos_freq, np.float32)
        sin = np.ascontiguousarray(rope_freqs.sin_freq, np.float32)
        probs_dtype = np.dtype(probs_dtype)
        if probs_dtype not in (np.dtype(np.float32), np.dtype(np.float16)):
            raise ValueError("probs_dtype must be float32 or float16")
        out_code = _lib.F16 if probs_dtype == np.dtype(np.float16) else _lib.F32
        inflight = []
        slot = 0

        def finish(item):
            s0, _keep, lg, pr = item
            _lib.check(eng.h, eng.L.a2m_collect_host(eng.h, s0), "a2m_collect_host", eng.L)
            if copy:
                return (None if lg is None else lg.copy()), pr.copy()
            return lg, pr

        for x in batches:
            if not (isinstance(x, np.ndarray) and x.dtype in (np.float32, np.float16) and x.flags.c_contiguous):
                x = np.ascontiguousarray(x, dtype=np.float32)
            if x.ndim != 3 or x.shape[1:] != (2, 80000):
                raise ValueError(f"batches must be (B, 2, 80000), got {x.shape}")
            if len(inflight) == HOST_SLOTS:
                yield finish(inflight.pop(0))
            B = x.shape[0]
            ring = self._out_ring.setdefault((B, want_logits, probs_dtype.str), {"bufs": [], "n": 0})
            if len(ring["bufs"]) < HOST_SLOTS + 1:
                ring["bufs"].append((pinned_empty((B, 250, 90)) if want_logits else None, pinned_empty((B, 250, 90), probs_dtype)))
                lg, pr = ring["bufs"][-1]
            else:
                lg, pr = ring["bufs"][ring["n"] % (HOST_SLOTS + 1)]
            ring["n"] += 1
            rc = eng.L.a2m_submit_host_ex(eng.h, slot, x.ctypes.data, _lib.F16 if x.dtype == np.float16 else _lib.F32, B,
                                          cos.ctypes.data, sin.ctypes.data, cos.shape[0],
                                          None if lg is None else lg.ctypes.data, pr.ctypes.data, out_code)
            _lib.check(eng.h, rc, "a2m_submit_host_ex", eng.L)
            inflight.append((slot, x, lg, pr))
            slot = (slot + 1) % HOST_SLOTS
        for item in inflight:
            yield finish(item)

    def profile_steps(self, batch: int, repeats: int = 5, device: Optional[int] = None):
        """Per-launch CUDA-event timings of the forward plan: list of (kernel, ms, algorithmic flops, bytes)."""
        eng = self._engine(_default_device() if device is None else device)
        n = eng.L.a2m_profile_steps(eng.h, batch, repeats, 0, None)
        if n < 0:
            _lib.check(eng.h, n, "a2m_profile_steps", eng.L)
        buf = (_lib.StepProfile * n)()
        m = eng.L.a2m_profile_steps(eng.h, batch, repeats, n, buf)
        if m < 0:
            _lib.check(eng.h, m, "a2m_profile_steps", eng.L)
        return [(buf[i].kernel.decode(), float(buf[i].ms), float(buf[i].flops), float(buf[i].bytes)) for i in range(m)]

    def last_launch_count(self, device: Optional[int] = None) -> int:
        eng = self._engine(_default_device() if device is None else device)
        return int(eng.L.a2m_last_launch_count(eng.h))


def fold_key(key) -> int:
    """PRNG key(s) -> one 64-bit dropout seed: an int, a jax-style uint32[2] key, or an array of per-sample keys (train.py:53)."""
    if key is None:
        return 0
    seed = 0
    for v in np.asarray(key).ravel().tolist():
        seed = ((seed ^ (int(v) & 0xFFFFFFFFFFFFFFFF)) * 0x9E3779B97F4A7C15 + 0x7F4A7C15) & 0xFFFFFFFFFFFFFFFF
    return seed


def _default_device() -> int:
    import os
    return int(os.environ.get("LOCAL_RANK", os.environ.get("A2M_DEVICE", "0")))


class _PinnedOwner:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            _lib.lib().a2m_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array backed by page-locked host memory (a2m_host_alloc): the host path copies to / from it directly,
    asynchronously, without the pageable -> pinned staging copy."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = _lib.lib().a2m_host_alloc(max(n, 1))
    if not ptr:
        raise MemoryError("a2m_host_alloc failed (is a CUDA device present?)")
    owner = _PinnedOwner(ptr)
    buf = (C.c_char * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[id(buf)] = owner  # keep the allocation alive as long as the ctypes buffer object
    import weakref
    weakref.finalize(buf, _PINNED.pop, id(buf), None)
    return arr


_PINNED: dict = {}


def vmap(fn, in_axes=(None, 0, None), out_axes=0, axis_name=None):
    """Adapter for the reference's two call shapes: jax.vmap(model.predict, in_axes=(None, 0, None)) (infer.py:40) and
    jax.vmap(model, in_axes=(0, None, None, 0, None), out_axes=(0, None), axis_name="batch") (train.py:56-58).  The kernels
    are natively batched, so the mapped axis is simply passed through."""
    axes = tuple(in_axes)
    if axes == (None, 0, None):
        def mapped(state, samples, rope_freqs):
            return fn(state, samples, rope_freqs)
        return mapped
    if axes == (0, None, None, 0, None):
        def mapped_train(samples, state, rope_freqs, keys, enable_dropout):
            return fn(samples, state, rope_freqs, keys, enable_dropout)
        return mapped_train
    raise NotImplementedError("supported: in_axes=(None, 0, None) for model.predict, (0, None, None, 0, None) for model.__call__")
