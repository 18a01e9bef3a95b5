"""Parameter pytree of the reference model, as plain nested dicts of numpy arrays.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Field names, nesting order and leaf shapes follow the reference's equinox
modules so that a flattened key path here ("layers.5.layers.3.gamma") is the
key path orbax would use for the same leaf:

* ``OutputSequenceGenerator`` fields  -- model.py:673-678
* ``Stem`` / ``Downsample``           -- model.py:84-118   (conv, norm)
* ``Block``                           -- model.py:120-126  (depth_conv, point_conv_1,
                                         point_conv_2, stochastic_depth_dropout, norm, gamma)
* ``TransformerStack.layers``         -- model.py:646-647  (every leaf stacked on a leading 8)
* ``TransformerLayer``                -- model.py:474-478
* ``SelfAttention``                   -- model.py:260-267  (bias-free Linear layers)
* ``LocalSelfAttention``              -- model.py:377-379  (wraps ``self_attention``)
* ``FeedForwardBlock``                -- model.py:200-203
* ``Decoder``                         -- model.py:169-171

Library semantics assumed: eqx.nn.Conv1d weight (out, in/groups, k),
bias (out, 1); eqx.nn.Linear weight (out, in), bias (out,); default init
U(-1/sqrt(fan_in), +1/sqrt(fan_in)) for weights and biases; LayerNorm weight 1, bias 0.
The random stream is numpy's PCG64, not JAX threefry: oracle and kernels share
the same arrays, so only the distribution matters.
"""
from __future__ import annotations

import numpy as np

MODEL_CONFIG = {  # model.py:20-34
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
VOCAB = 90            # audio_to_midi_dataset.py:26  MIDI_EVENT_VOCCAB_SIZE
WINDOW_SECONDS = 5.0  # audio_to_midi_dataset.py:28  MODEL_AUDIO_LENGTH
SAMPLE_RATE = 16000   # audio_to_midi_dataset.py:111
WINDOW_SAMPLES = 80000
LOCAL_CONTEXT = 16    # model.py:635
EXPECTED_PARAM_COUNT = 11_606_269


def _uniform(rng, shape, fan_in):
    lim = 1.0 / np.sqrt(fan_in)
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def _conv(rng, cout, cin_per_group, k):
    fan_in = cin_per_group * k
    return {"weight": _uniform(rng, (cout, cin_per_group, k), fan_in),
            "bias": _uniform(rng, (cout, 1), fan_in)}


def _linear(rng, cout, cin, bias=True, lead=()):
    d = {"weight": _uniform(rng, lead + (cout, cin), cin)}
    if bias:
        d["bias"] = _uniform(rng, lead + (cout,), cin)
    return d


def _ln(n, lead=()):
    return {"weight": np.ones(lead + (n,), np.float32), "bias": np.zeros(lead + (n,), np.float32)}


def _attention(rng, d, heads, hd, ckv, lead):
    return {
        "query_up_proj": _linear(rng, heads * hd, d, bias=False, lead=lead),
        "kv_down_proj": _linear(rng, ckv, d, bias=False, lead=lead),
        "key_up_proj": _linear(rng, heads * hd, ckv, bias=False, lead=lead),
        "value_up_proj": _linear(rng, heads * hd, ckv, bias=False, lead=lead),
        "output_proj": _linear(rng, d, heads * hd, bias=False, lead=lead),
    }


def _transformer_layer(rng, d, heads, hd, ckv, inter, lead, local):
    att = _attention(rng, d, heads, hd, ckv, lead)
    return {
        "attention_norm": _ln(d, lead),
        "attention_block": {"self_attention": att} if local else att,
        "feed_forward_norm": _ln(d, lead),
        "feed_forward_block": {
            "attention_to_intermediate_proj": _linear(rng, 2 * inter, d, lead=lead),
            "intermediate_to_attention_proj": _linear(rng, d, inter, lead=lead),
        },
    }


def init_params(seed: int = 1234, conf: dict | None = None, gamma_mode: str = "default",
                decoder_gain: float = 1.0, trained_like: bool = False) -> dict:
    """Random-init parameter tree.

    gamma_mode "default" -> layer scale 1e-6 (model.py:157-158); "active" -> U(0.5, 1.5)
    so that the inside of every Block matters to the output (SURVEY.md parity trap 2).
    decoder_gain multiplies the decoder weight so probabilities spread over (0, 1).
    trained_like perturbs every LayerNorm weight/bias away from 1/0 so affine terms are tested.
    """
    conf = dict(MODEL_CONFIG if conf is None else conf)
    rng = np.random.Generator(np.random.PCG64(seed))
    dims, depths = conf["dims"], conf["depths"]
    hidden = [int(d * conf["cnn_hidden_expansion"]) for d in dims]
    sdd = np.linspace(0.0, conf["sdd_rate"], sum(depths)).astype(np.float32)  # model.py:694

    stages, depth_count = [], 0
    for i, c in enumerate(dims):
        seq = []
        if i == 0:
            seq.append({"conv": _conv(rng, c, 2, 5), "norm": _ln(c)})            # Stem
        else:
            seq.append({"conv": _conv(rng, c, dims[i - 1], 2), "norm": _ln(dims[i - 1])})  # Downsample
        for j in range(depths[i]):
            if gamma_mode == "default":
                gamma = np.full((c,), 1e-6, np.float32)
            elif gamma_mode == "active":
                gamma = rng.uniform(0.5, 1.5, size=(c,)).astype(np.float32)
            else:
                raise ValueError(gamma_mode)
            seq.append({
                "depth_conv": _conv(rng, c, 1, 7),
                "point_conv_1": _conv(rng, hidden[i], c, 1),
                "point_conv_2": _conv(rng, c, hidden[i], 1),
                "stochastic_depth_dropout": {"p": np.float32(sdd[depth_count + j])},
                "norm": _ln(c),
                "gamma": gamma,
            })
        depth_count += depths[i]
        stages.append({"layers": seq})

    d = conf.get("transformer_hidden_dim", dims[-1])
    assert d == dims[-1], "transformer_projection is None in the default config (model.py:718-724)"
    nl = conf["num_transformer_layers"]
    heads, hd = conf["num_transformer_heads"], conf["attention_size"]
    ckv = conf["compressed_attention_kv_size"]
    inter = int(d * conf["transformer_hidden_expansion"])
    lead = (nl,)
    tree = {
        "layers": stages,
        "norm": _ln(dims[-1]),
        "transformer": {"layers": {
            "local_attention": _transformer_layer(rng, d, heads, hd, ckv, inter, lead, local=True),
            "global_attention": _transformer_layer(rng, d, heads, hd, ckv, inter, lead, local=False),
        }},
        "decoder": {"decoder_pooling": _linear(rng, VOCAB, d), "norm": _ln(d)},
    }
    if decoder_gain != 1.0:
        tree["decoder"]["decoder_pooling"]["weight"] *= np.float32(decoder_gain)
    if trained_like:
        for path, leaf in flatten(tree).items():
            if ".norm." in "." + path or path.startswith("norm.") or "_norm." in path:
                if path.endswith("weight"):
                    leaf *= rng.uniform(0.7, 1.3, size=leaf.shape).astype(np.float32)
                else:
                    leaf += rng.uniform(-0.2, 0.2, size=leaf.shape).astype(np.float32)
    return tree


def flatten(tree, prefix="") -> dict:
    """Depth-first, insertion-ordered {dotted key path: leaf}."""
    out = {}
    if isinstance(tree, dict):
        for k, v in tree.items():
            out.update(flatten(v, f"{prefix}{k}."))
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            out.update(flatten(v, f"{prefix}{i}."))
    else:
        out[prefix[:-1]] = tree
    return out


def param_count(tree) -> int:
    return int(sum(np.asarray(v).size for v in flatten(tree).values()))


def cast(tree, dtype):
    if isinstance(tree, dict):
        return {k: cast(v, dtype) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return [cast(v, dtype) for v in tree]
    return np.asarray(tree, dtype=dtype)


def layer_slice(tree, i):
    """Select transformer layer i from leaves stacked on a leading axis (what lax.scan does, model.py:668)."""
    if isinstance(tree, dict):
        return {k: layer_slice(v, i) for k, v in tree.items()}
    return tree[i]
