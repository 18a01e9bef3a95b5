"""Shared by tests/test_gpu_train.py and tools/check_grads.py: CUDA gradients (through the C ABI) vs torch-CPU autograd
of the oracle (oracle/model_torch.py loss_fn, train.py:39-62 with dropout off)."""
import numpy as np
import torch

import audio_to_midi_b200 as A
from audio_to_midi_b200 import train as T
from oracle import model_torch as MT
from oracle import params as P
from oracle import synth


def oracle_grads(tree, audio, labels, scale=1.0, masks=None):
    tp = MT.to_torch(tree, requires_grad=False)
    leaves = {}

    def mark(t, prefix=""):
        if isinstance(t, dict):
            return {k: mark(v, f"{prefix}{k}.") for k, v in t.items()}
        if isinstance(t, list):
            return [mark(v, f"{prefix}{i}.") for i, v in enumerate(t)]
        if t.dtype.is_floating_point and t.ndim >= 0:
            t = t.clone().requires_grad_(True)
            leaves[prefix[:-1]] = t
        return t

    tp = mark(tp)
    loss, logits = MT.loss_fn(tp, torch.tensor(audio), torch.tensor(labels), scale=scale, masks=masks)
    loss.backward()
    return float(loss), {k: (v.grad.numpy() if v.grad is not None else None) for k, v in leaves.items()}, logits.detach().numpy()


def cuda_grads(tree, audio, labels, scale=1.0, device=0, dropout=0.0, seed=0):
    model = A.OutputSequenceGenerator(A.model_config, key=0).load_leaves(P.flatten(tree))
    eng = T.TrainEngine(model, device)
    rope = A.precompute_frequencies(64, 300)
    dev = torch.device(f"cuda:{device}")
    eng.zero_grad()
    eng.set_dropout(dropout, seed)
    logits = eng.forward_backward(torch.tensor(audio, device=dev), torch.tensor(labels, device=dev), rope, scale=scale, want_logits=True)
    torch.cuda.synchronize()
    return float(eng.loss.item()), eng.grads_tree(), logits.cpu().numpy(), eng


def compare(gref, gcuda):
    """Per-leaf relative L2 error, sorted worst first: list of (path, rel, |ref|, |cuda|)."""
    rows = []
    for k, r in gref.items():
        if r is None or k.endswith("stochastic_depth_dropout.p"):
            continue
        c = gcuda[k]
        nr, nc = float(np.linalg.norm(r)), float(np.linalg.norm(c))
        rel = float(np.linalg.norm(c.astype(np.float64) - r.astype(np.float64)) / max(nr, 1e-30))
        rows.append((k, rel, nr, nc))
    rows.sort(key=lambda x: -x[1])
    return rows


def setup(batch=2, seed=7):
    tree = P.init_params(seed, gamma_mode="active", decoder_gain=4.0, trained_like=True)
    audio, labels = synth.make_windows(batch, seed, with_labels=True)
    return tree, audio, labels
