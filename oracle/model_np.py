"""NumPy restatement (fp64 by default) of the reference forward for ONE window.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (no reference fixtures exist).

Every function cites the reference lines it follows.  The model is unbatched,
exactly like the reference (callers ``vmap``; here: a Python loop in ``forward_batch``).
It is written "as the reference is written" -- local attention really gathers 31
windows, projects each of them and scatter-adds with the reference's padded
indices -- so that the de-duplicated algebra used by the CUDA kernels is checked
against the literal algorithm, not against itself.
"""
from __future__ import annotations

import numpy as np

from .params import LOCAL_CONTEXT, MODEL_CONFIG, layer_slice

LN_EPS = 1e-5  # eqx.nn.LayerNorm default


# --------------------------------------------------------------------------- primitives
def layer_norm(x, p, axis=-1):
    """eqx.nn.LayerNorm over ``axis``: biased variance, eps inside the sqrt, then affine."""
    mean = x.mean(axis=axis, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=axis, keepdims=True)
    y = (x - mean) / np.sqrt(var + LN_EPS)
    shape = [1] * x.ndim
    shape[axis] = -1
    return y * p["weight"].reshape(shape) + p["bias"].reshape(shape)


def gelu_tanh(x):
    """jax.nn.gelu default (approximate=True)."""
    return 0.5 * x * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * x ** 3)))


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def softmax_last(x):
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


def conv1d_strided(x, p, k):
    """eqx.nn.Conv1d(kernel=k, stride=k, VALID) on (Cin, L) -> (Cout, L // k); cross-correlation."""
    cin, L = x.shape
    lo = L // k
    xr = x[:, : lo * k].reshape(cin, lo, k)                 # (Cin, Lo, k)
    return np.einsum("ock,clk->ol", p["weight"], xr) + p["bias"]


def depthwise_conv7_same(x, p):
    """eqx.nn.Conv1d(groups=C, kernel=7, padding='SAME') -> zero pad (3, 3); cross-correlation."""
    c, L = x.shape
    xp = np.pad(x, ((0, 0), (3, 3)))
    out = np.zeros_like(x)
    w = p["weight"][:, 0, :]                                 # (C, 7)
    for t in range(7):
        out += w[:, t:t + 1] * xp[:, t:t + L]
    return out + p["bias"]


def pointwise(x, p):
    """eqx.nn.Conv1d(kernel=1)."""
    return p["weight"][:, :, 0] @ x + p["bias"]


def linear(x, p):
    """vmap(eqx.nn.Linear) over rows of x (S, in) -> (S, out)."""
    y = x @ p["weight"].T
    if "bias" in p:
        y = y + p["bias"]
    return y


# --------------------------------------------------------------------------- CNN frontend
def stem(x, p):
    """model.py:98-100."""
    return layer_norm(conv1d_strided(x, p["conv"], 5), p["norm"], axis=0)


def downsample(x, p):
    """model.py:116-118: LN over the INPUT channels, then conv k2 s2."""
    return conv1d_strided(layer_norm(x, p["norm"], axis=0), p["conv"], 2)


def block(x, p):
    """model.py:160-167 with enable_dropout=False (Sequential never forwards it; SDD is the identity)."""
    out = depthwise_conv7_same(x, p["depth_conv"])
    out = layer_norm(out, p["norm"], axis=0)
    out = pointwise(out, p["point_conv_1"])
    out = gelu_tanh(out)
    out = pointwise(out, p["point_conv_2"])
    out = p["gamma"][:, None] * out
    return out + x


def cnn_frontend(samples, params, taps=None):
    """model.py:756-762 -> (T, D)."""
    h = samples
    for si, stage in enumerate(params["layers"]):
        seq = stage["layers"]
        h = stem(h, seq[0]) if si == 0 else downsample(h, seq[0])
        for blk in seq[1:]:
            h = block(h, blk)
        if taps is not None:
            taps[f"stage{si}"] = h.T.copy()
    h = layer_norm(h, params["norm"], axis=0)
    return h.T


# --------------------------------------------------------------------------- RoPE
def precompute_frequencies(dim, max_pos, theta=10000.0, dtype=np.float64):
    """rope.py:12-22.  The reference builds the table in fp32 (inv_freq, outer product, cos/sin all
    fp32); it is an INPUT of the model, so the oracle keeps the fp32 table and only widens it."""
    f32 = np.float32
    inv_freq = f32(1.0) / (f32(theta) ** (np.arange(0, dim, 2, dtype=f32)[: dim // 2] / f32(dim)))
    t = np.arange(0, max_pos, dtype=f32)
    freqs = np.outer(t, inv_freq).astype(f32)
    return np.cos(freqs).astype(dtype), np.sin(freqs).astype(dtype)


def calculate_rope(x, rope):
    """rope.py:25-53.  x is (S, heads, head_dim); positions are 0..S-1 (table sliced to [:S])."""
    cos_t, sin_t = rope
    s = x.shape[0]
    cos = cos_t[:s, None, :]
    sin = sin_t[:s, None, :]
    x1 = x[..., 0::2]
    x2 = x[..., 1::2]
    out = np.stack([x1 * cos - x2 * sin, x1 * sin + x2 * cos], axis=-1)
    return out.reshape(x.shape)


# --------------------------------------------------------------------------- transformer
def dot_product_attention(q, k, v):
    """model.py:241-257, dropout off."""
    q = q / np.sqrt(q.shape[-1])
    logits = q @ k.T
    w = softmax_last(logits)
    return w @ v


def self_attention(x, p, rope, num_heads):
    """model.py:340-374."""
    s = x.shape[0]
    q = calculate_rope(linear(x, p["query_up_proj"]).reshape(s, num_heads, -1), rope)
    c_kv = linear(x, p["kv_down_proj"])
    k = calculate_rope(linear(c_kv, p["key_up_proj"]).reshape(s, num_heads, -1), rope)
    v = linear(c_kv, p["value_up_proj"]).reshape(s, num_heads, -1)
    heads = [dot_product_attention(q[:, h], k[:, h], v[:, h]) for h in range(num_heads)]
    attn = np.stack(heads, axis=1).reshape(s, -1)
    return linear(attn, p["output_proj"])


def local_self_attention(x, p, rope, num_heads, window=LOCAL_CONTEXT):
    """model.py:409-471, literally: pad, gather windows, attend, scatter-add with PADDED indices
    into an UNPADDED buffer (out-of-range updates dropped, JAX default), divide by count."""
    seq_len, hidden = x.shape
    stride = window // 2
    required = stride - (seq_len - window) % stride
    xin = x
    if required != stride:
        if required % 2 == 0:
            xin = np.pad(x, ((required // 2, required // 2), (0, 0)))
        else:
            xin = np.pad(x, ((required // 2, required // 2 + 1), (0, 0)))
    num_windows = (xin.shape[0] - window) // stride + 1
    out = np.zeros((seq_len, hidden), dtype=x.dtype)
    count = np.zeros((seq_len,), dtype=x.dtype)
    for w in range(num_windows):
        start = w * stride
        ow = self_attention(xin[start:start + window], p["self_attention"], rope, num_heads)
        for t in range(window):
            idx = start + t
            if idx < seq_len:              # .at[idx].add drops out-of-bounds updates
                out[idx] += ow[t]
                count[idx] += 1
    return out / count[:, None]


def feed_forward(x, p):
    """model.py:226-238, dropout off."""
    u = linear(x, p["attention_to_intermediate_proj"])
    x1, x2 = np.split(u, 2, axis=-1)
    h = gelu_tanh(x1) * x2
    return linear(h, p["intermediate_to_attention_proj"])


def transformer_layer(x, p, rope, num_heads, local):
    """model.py:529-556, dropout off."""
    n = layer_norm(x, p["attention_norm"])
    if local:
        r = local_self_attention(n, p["attention_block"], rope, num_heads)
    else:
        r = self_attention(n, p["attention_block"], rope, num_heads)
    h = x + r
    return h + feed_forward(layer_norm(h, p["feed_forward_norm"]), p["feed_forward_block"])


def transformer_stack(x, p, rope, num_heads, num_layers, taps=None):
    """model.py:649-670 + 599-612: scan over stacked layers; each = local then global."""
    for i in range(num_layers):
        lp = layer_slice(p["layers"], i)
        x = transformer_layer(x, lp["local_attention"], rope, num_heads, local=True)
        if taps is not None:
            taps[f"tl{i}_local"] = x.copy()
        x = transformer_layer(x, lp["global_attention"], rope, num_heads, local=False)
        if taps is not None:
            taps[f"tl{i}_global"] = x.copy()
    return x


def decoder(x, p):
    """model.py:185-198."""
    logits = linear(layer_norm(x, p["norm"]), p["decoder_pooling"])
    return logits, sigmoid(logits)


# --------------------------------------------------------------------------- whole model
def forward(params, samples, rope=None, conf=None, taps=None):
    """OutputSequenceGenerator.__call__ (model.py:740-769), key=None, enable_dropout=False.

    samples (2, N) -> (logits (T, 90), probs (T, 90)).  Computation dtype = samples.dtype of
    the caller-cast params (pass fp64 arrays for the reference-quality result).
    """
    conf = MODEL_CONFIG if conf is None else conf
    if rope is None:
        rope = precompute_frequencies(conf["attention_size"], 300, dtype=samples.dtype)  # infer.py:38
    h = cnn_frontend(samples, params, taps)
    if taps is not None:
        taps["cnn_out"] = h.copy()
    h = transformer_stack(h, params["transformer"], rope, conf["num_transformer_heads"],
                          conf["num_transformer_layers"], taps)
    return decoder(h, params["decoder"])


def forward_batch(params, samples, conf=None, dtype=np.float64):
    """What jax.vmap(model.predict, in_axes=(None, 0, None)) computes (infer.py:40)."""
    from .params import cast
    p = cast(params, dtype)
    conf = MODEL_CONFIG if conf is None else conf
    rope = precompute_frequencies(conf["attention_size"], 300, dtype=dtype)
    outs = [forward(p, np.asarray(s, dtype=dtype), rope, conf) for s in samples]
    return np.stack([o[0] for o in outs]), np.stack([o[1] for o in outs])


def bce_with_logits_sum(logits, targets, scale=1.0):
    """train.py:39-47: optax.sigmoid_binary_cross_entropy summed over (T, 90), times scale; per sample."""
    z = logits
    log_sig = -np.logaddexp(0.0, -z)       # log sigmoid(z)
    log_nsig = -np.logaddexp(0.0, z)       # log sigmoid(-z)
    loss = -targets * log_sig - (1.0 - targets) * log_nsig
    return (loss * scale).sum(axis=(-2, -1))
