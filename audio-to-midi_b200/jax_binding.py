"""JAX side of the boundary: jax.ffi custom calls into liba2m_xla_ffi.so (csrc/a2m_xla_ffi.cc), which forwards
to the plain C ABI of include/a2m.h.  UNTESTED IN THIS IMAGE -- jax / jaxlib are not installed, so importing
this module raises ImportError here; it is what a reference maintainer drops next to infer.py / train.py where JAX exists.

    import audio_to_midi_b200.jax_binding as jb
    b = jb.Binding(model_leaves)                # {pytree key path: np.ndarray}; a2m_create_ex + a2m_train_init per local device
    logits, probs = b.predict(samples, rope)    # replaces jax.vmap(model.predict, (None, 0, None))(state, samples, rope)   infer.py:40
    (loss, grads) = b.loss_and_grad(params_blob, audio, events, rope, scale, key)     # replaces compute_loss' value_and_grad   train.py:48-62
    logits = b.model_logits(params_blob, audio, rope, key)   # custom_vjp pair for any other loss on the logits
"""
from __future__ import annotations

import ctypes
import functools
import os

import jax            # noqa: F401  (ImportError here is the documented state of this image)
import jax.numpy as jnp
import numpy as np

from . import _lib
from .build import build_xla_ffi
from .model import _Engine, fold_key, model_config

_TARGETS = {"a2m_forward": "A2mForward", "a2m_forward_train": "A2mForwardTrain", "a2m_backward": "A2mBackward",
            "a2m_loss_and_grad": "A2mLossAndGrad", "a2m_allreduce": "A2mAllReduce", "a2m_adamw": "A2mAdamW"}
_registered = False


def register():
    global _registered
    if _registered:
        return
    _lib.lib()                                           # libaudio2midi_b200.so first: the shim links against it
    shim = ctypes.CDLL(build_xla_ffi(), mode=ctypes.RTLD_GLOBAL)
    for name, sym in _TARGETS.items():
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(shim, sym)), platform="CUDA")
    _registered = True


class Binding:
    """One A2mHandle per local device, created eagerly (handles cannot be created from inside a traced computation)."""

    def __init__(self, leaves: dict, device: int = 0, train: bool = True):
        register()
        self.eng = _Engine(device)
        self.paths = list(leaves)
        blob, table, self.offsets = _Engine.blob_and_table([(p, np.asarray(leaves[p], np.float32)) for p in self.paths])
        fn = self.eng.L.a2m_train_init if train else self.eng.L.a2m_load_weights
        _lib.check(self.eng.h, fn(self.eng.h, blob.ctypes.data, blob.nbytes, table, len(table)), fn.__name__)
        self.n_params = blob.size
        self.handle = np.int64(self.eng.h.value)
        self.params_blob = jnp.asarray(blob)

    # ---- infer.py:40
    def predict(self, samples, rope_freqs):
        B = samples.shape[0]
        out = (jax.ShapeDtypeStruct((B, 250, 90), jnp.float32),) * 2
        return jax.ffi.ffi_call("a2m_forward", out, vmap_method="sequential")(
            samples.astype(jnp.float32), rope_freqs.cos_freq, rope_freqs.sin_freq, handle=self.handle)

    # ---- train.py:48-62 in one call
    def loss_and_grad(self, params_blob, audio, events, rope_freqs, scale, key, dropout_rate=None):
        rate = model_config["transformer_dropout_rate"] if dropout_rate is None else dropout_rate
        out = (jax.ShapeDtypeStruct((1,), jnp.float32), jax.ShapeDtypeStruct((self.n_params,), jnp.float32))
        loss, grads = jax.ffi.ffi_call("a2m_loss_and_grad", out)(
            params_blob, audio.astype(jnp.float32), events.astype(jnp.float32), rope_freqs.cos_freq, rope_freqs.sin_freq,
            handle=self.handle, scale=np.float32(scale), dropout_rate=np.float32(rate), seed=np.int64(fold_key(key) & (2 ** 63 - 1)))
        return loss[0], grads

    # ---- custom_vjp: logits as a differentiable function of the parameter blob (any loss on top)
    def model_logits(self, params_blob, audio, rope_freqs, key, dropout_rate=None):
        rate = np.float32(model_config["transformer_dropout_rate"] if dropout_rate is None else dropout_rate)
        seed = np.int64(fold_key(key) & (2 ** 63 - 1))
        B = audio.shape[0]

        @jax.custom_vjp
        def f(p, a):
            return jax.ffi.ffi_call("a2m_forward_train", jax.ShapeDtypeStruct((B, 250, 90), jnp.float32))(
                p, a, rope_freqs.cos_freq, rope_freqs.sin_freq, handle=self.handle, dropout_rate=rate, seed=seed)

        def fwd(p, a):
            return f(p, a), None                       # the tape stays inside the handle

        def bwd(_, dlogits):
            g = jax.ffi.ffi_call("a2m_backward", jax.ShapeDtypeStruct((self.n_params,), jnp.float32))(dlogits, handle=self.handle)
            return g, None                             # no gradient with respect to the audio (the reference takes none either)

        f.defvjp(fwd, bwd)
        return f(params_blob, audio.astype(jnp.float32))

    # ---- train.py:324-325
    def adamw(self, grads, lr, step, cfg, grad_divisor=1.0):
        out = (jax.ShapeDtypeStruct((self.n_params,), jnp.float32), jax.ShapeDtypeStruct((2,), jnp.float32))
        return jax.ffi.ffi_call("a2m_adamw", out)(
            grads, handle=self.handle, lr=np.float32(lr), b1=np.float32(cfg.b1), b2=np.float32(cfg.b2), eps=np.float32(cfg.eps),
            weight_decay=np.float32(cfg.weight_decay), grad_divisor=np.float32(grad_divisor), clip_norm=np.float32(cfg.clip_norm),
            step=np.int32(step))

    def allreduce(self, grads):
        return jax.ffi.ffi_call("a2m_allreduce", jax.ShapeDtypeStruct((self.n_params,), jnp.float32), input_output_aliases={0: 0})(
            grads, handle=self.handle)
