"""PyTorch-CPU twin of the reference forward, batched, differentiable.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.

Independent of oracle/model_np.py on purpose (library convs / layer_norm / softmax instead
of hand-rolled einsums) so the two restatements pin each other.  It also serves as
  * the fp32 "CPU restatement (JAX absent)" timed by bench.py's cpu_baseline / --impl reference,
  * the autograd reference for the backward kernels (train.py:39-62).
It follows the reference as written: local attention projects all 31 x 16 windowed tokens
(model.py:439-449) rather than the 250 unique ones.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .params import LOCAL_CONTEXT, MODEL_CONFIG


def to_torch(tree, dtype=torch.float32, requires_grad=False):
    if isinstance(tree, dict):
        return {k: to_torch(v, dtype, requires_grad) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return [to_torch(v, dtype, requires_grad) for v in tree]
    t = torch.tensor(np.asarray(tree), dtype=dtype)
    if requires_grad:
        t.requires_grad_(True)
    return t


def _ln_channels(x, p):
    """LayerNorm over the channel axis of (B, C, L)  (model.py:100,117,162,759)."""
    c = x.shape[1]
    return F.layer_norm(x.transpose(1, 2), (c,), p["weight"], p["bias"], 1e-5).transpose(1, 2)


def _at_least_f32(x):
    """The reference's `.astype(jnp.float32)` upcasts (from fp16); never a DOWNcast when the twin runs in fp64."""
    return x.to(torch.promote_types(x.dtype, torch.float32))


def _gelu(x):
    return F.gelu(x, approximate="tanh")


def precompute_frequencies(dim, max_pos, theta=10000.0):
    """rope.py:12-22 (fp32)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    t = torch.arange(0, max_pos, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    return torch.cos(freqs), torch.sin(freqs)


def _rope(x, rope):
    """rope.py:25-53; x (..., S, H, hd), position = index along S."""
    cos_t, sin_t = rope
    s = x.shape[-3]
    cos = cos_t[:s, None, :].to(x.dtype)
    sin = sin_t[:s, None, :].to(x.dtype)
    x1, x2 = x[..., 0::2], x[..., 1::2]
    return torch.stack([x1 * cos - x2 * sin, x1 * sin + x2 * cos], dim=-1).flatten(-2)


def _mix32(x):
    """numpy restatement of the CUDA path's counter-based dropout hash (csrc/ptx.cuh mix32)."""
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7feb352d)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846ca68b)
    x ^= x >> np.uint32(16)
    return x


def dropout_masks(seed: int, rate: float, batch: int, num_layers: int = 8):
    """The masks (already scaled by 1 / (1 - rate)) the CUDA training path applies for (seed, rate): a pure function of
    (seed, site, element index), csrc/ptx.cuh.  Returns {("ffn" | "global" | "local", layer): torch.Tensor}; layer counts
    TransformerLayers 0..15 (even = local, odd = global)."""
    p = min(max(float(rate), 0.0), 0.999)
    thresh = np.uint32(min(int(p * 4294967296.0), 4294967295))
    inv_keep = np.float32(1.0 / (1.0 - p))
    seed32 = np.uint32((seed & 0xffffffff) ^ (((seed >> 32) * 0x9e3779b9) & 0xffffffff))

    def mul(site, idx):
        key = _mix32(np.array([seed32 ^ np.uint32((site * 0x85ebca6b + 0x9e3779b9) & 0xffffffff)], np.uint32))[0]
        h = _mix32((idx.astype(np.uint32) * np.uint32(0x9e3779b1)) ^ key)
        return torch.tensor(np.where(h >= thresh, inv_keep, np.float32(0.0)).astype(np.float32))

    out = {}
    with np.errstate(over="ignore"):
        b = np.arange(batch, dtype=np.int64)
        for i in range(2 * num_layers):
            idx = ((b[:, None, None] * 256 + np.arange(250)[None, :, None]) * 256 + np.arange(256)[None, None, :])
            out[("ffn", i)] = mul(i * 4 + 0, idx)
            bh = b[:, None] * 4 + np.arange(4)[None, :]
            if i % 2 == 1:
                idx = ((bh[:, :, None, None] * 256 + np.arange(250)[None, None, :, None]) * 256 + np.arange(250)[None, None, None, :])
                out[("global", i)] = mul(i * 4 + 1, idx)
            else:
                idx = (((bh[:, :, None, None, None] * 31 + np.arange(31)[None, None, :, None, None]) * 16
                        + np.arange(16)[None, None, None, :, None]) * 16 + np.arange(16)[None, None, None, None, :])
                out[("local", i)] = mul(i * 4 + 2, idx).permute(0, 2, 1, 3, 4)   # (B, window, head, row, key)
    return out


def _self_attention(x, p, rope, heads, dropout_p=0.0, wmask=None):
    """model.py:340-374 on (..., S, D)."""
    s = x.shape[-2]
    lead = x.shape[:-2]
    q = _rope(F.linear(x, p["query_up_proj"]["weight"]).reshape(*lead, s, heads, -1), rope)
    c = F.linear(x, p["kv_down_proj"]["weight"])
    k = _rope(F.linear(c, p["key_up_proj"]["weight"]).reshape(*lead, s, heads, -1), rope)
    v = F.linear(c, p["value_up_proj"]["weight"]).reshape(*lead, s, heads, -1)
    q, k, v = (t.transpose(-3, -2) for t in (q, k, v))              # (..., H, S, hd)
    logits = (q / math.sqrt(q.shape[-1])) @ k.transpose(-1, -2)
    w = torch.softmax(_at_least_f32(logits), dim=-1).to(logits.dtype)     # model.py:252 upcasts to fp32
    if wmask is not None:
        w = w * wmask
    elif dropout_p > 0.0:
        w = F.dropout(w, dropout_p, training=True)
    a = (w @ v).transpose(-3, -2).reshape(*lead, s, -1)
    return F.linear(a, p["output_proj"]["weight"])


def _local_self_attention(x, p, rope, heads, window=LOCAL_CONTEXT, dropout_p=0.0, wmask=None):
    """model.py:409-471 on (B, S, D)."""
    b, seq_len, d = x.shape
    stride = window // 2
    required = stride - (seq_len - window) % stride
    xin = x
    if required != stride:
        lo = required // 2
        hi = required // 2 + (required % 2)
        xin = F.pad(x, (0, 0, lo, hi))
    nw = (xin.shape[1] - window) // stride + 1
    wins = xin.unfold(1, window, stride).permute(0, 1, 3, 2)         # (B, nw, window, D)
    ow = _self_attention(wins, p["self_attention"], rope, heads, dropout_p, wmask)
    idx = (torch.arange(nw)[:, None] * stride + torch.arange(window)[None, :]).reshape(-1)
    keep = idx < seq_len                                             # out-of-range scatter updates dropped
    flat = ow.reshape(b, nw * window, d)
    out = torch.zeros(b, seq_len, d, dtype=x.dtype).index_add(1, idx[keep], flat[:, keep])
    count = torch.zeros(seq_len, dtype=x.dtype).index_add(0, idx[keep], torch.ones(int(keep.sum()), dtype=x.dtype))
    return out / count[None, :, None]


def _feed_forward(x, p, dropout_p=0.0, omask=None):
    u = F.linear(x, p["attention_to_intermediate_proj"]["weight"], p["attention_to_intermediate_proj"]["bias"])
    x1, x2 = u.chunk(2, dim=-1)
    y = F.linear(_gelu(x1) * x2, p["intermediate_to_attention_proj"]["weight"],
                 p["intermediate_to_attention_proj"]["bias"])
    if omask is not None:
        y = y * omask
    elif dropout_p > 0.0:
        y = F.dropout(y, dropout_p, training=True)
    return y


def _transformer_layer(x, p, rope, heads, local, dropout_p=0.0, wmask=None, omask=None):
    d = x.shape[-1]
    n = F.layer_norm(x, (d,), p["attention_norm"]["weight"], p["attention_norm"]["bias"], 1e-5)
    r = (_local_self_attention(n, p["attention_block"], rope, heads, dropout_p=dropout_p, wmask=wmask) if local
         else _self_attention(n, p["attention_block"], rope, heads, dropout_p, wmask))
    h = x + r
    n2 = F.layer_norm(h, (d,), p["feed_forward_norm"]["weight"], p["feed_forward_norm"]["bias"], 1e-5)
    return h + _feed_forward(n2, p["feed_forward_block"], dropout_p, omask)


def _slice(tree, i):
    if isinstance(tree, dict):
        return {k: _slice(v, i) for k, v in tree.items()}
    return tree[i]


def forward(params, samples, rope=None, conf=None, dropout_p=0.0, taps=None, masks=None):
    """Batched OutputSequenceGenerator forward: samples (B, 2, N) -> logits, probs (B, T, 90)."""
    conf = MODEL_CONFIG if conf is None else conf
    if rope is None:
        rope = precompute_frequencies(conf["attention_size"], 300)
    h = samples
    for si, stage in enumerate(params["layers"]):
        seq = stage["layers"]
        if si == 0:
            h = _ln_channels(F.conv1d(h, seq[0]["conv"]["weight"], seq[0]["conv"]["bias"][:, 0], stride=5),
                             seq[0]["norm"])
        else:
            h = F.conv1d(_ln_channels(h, seq[0]["norm"]), seq[0]["conv"]["weight"],
                         seq[0]["conv"]["bias"][:, 0], stride=2)
        for blk in seq[1:]:
            c = h.shape[1]
            o = F.conv1d(h, blk["depth_conv"]["weight"], blk["depth_conv"]["bias"][:, 0], padding=3, groups=c)
            o = _ln_channels(o, blk["norm"])
            o = F.conv1d(o, blk["point_conv_1"]["weight"], blk["point_conv_1"]["bias"][:, 0])
            o = _gelu(o)
            o = F.conv1d(o, blk["point_conv_2"]["weight"], blk["point_conv_2"]["bias"][:, 0])
            h = blk["gamma"][None, :, None] * o + h
        if taps is not None:
            taps[f"stage{si}"] = h.transpose(1, 2)
    h = _ln_channels(h, params["norm"]).transpose(1, 2)                 # (B, T, D)
    if taps is not None:
        taps["cnn_out"] = h
    heads = conf["num_transformer_heads"]
    tl = params["transformer"]["layers"]
    for i in range(conf["num_transformer_layers"]):
        lp = _slice(tl, i)
        mk = masks or {}
        h = _transformer_layer(h, lp["local_attention"], rope, heads, True, dropout_p, mk.get(("local", 2 * i)), mk.get(("ffn", 2 * i)))
        if taps is not None:
            taps[f"tl{i}_local"] = h
        h = _transformer_layer(h, lp["global_attention"], rope, heads, False, dropout_p, mk.get(("global", 2 * i + 1)),
                               mk.get(("ffn", 2 * i + 1)))
        if taps is not None:
            taps[f"tl{i}_global"] = h
    dec = params["decoder"]
    n = F.layer_norm(h, (h.shape[-1],), dec["norm"]["weight"], dec["norm"]["bias"], 1e-5)
    logits = F.linear(n, dec["decoder_pooling"]["weight"], dec["decoder_pooling"]["bias"])
    return logits, torch.sigmoid(logits)


def loss_fn(params, samples, targets, scale=1.0, rope=None, conf=None, masks=None):
    """train.py:39-62: per-sample sum of BCE-with-logits x scale, mean over batch.  Dropout is off unless `masks`
    (dropout_masks) replays the CUDA path's masks."""
    logits, _ = forward(params, samples, rope, conf, masks=masks)
    per = F.binary_cross_entropy_with_logits(_at_least_f32(logits), targets, reduction="none").sum(dim=(1, 2)) * scale
    return per.mean(), logits
