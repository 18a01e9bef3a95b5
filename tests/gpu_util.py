"""Helpers for the -m gpu tests: device buffers via torch, calls through the C ABI (ctypes)."""
import ctypes as C

import numpy as np
import torch

import audio_to_midi_b200 as A
from audio_to_midi_b200 import _lib
from oracle import params as P


def make_model(seed, precision=None, **kw):
    """Product model carrying exactly the oracle's parameter arrays (precision: "bf16" | "f16" operand variant, None = default)."""
    tree = P.init_params(seed, **kw)
    m = A.OutputSequenceGenerator(A.model_config, key=0)
    m.load_leaves(P.flatten(tree))
    if precision is not None:
        m.precision = precision
    return m, tree


def engine(model, device=0):
    return model._engine(device)


def rope_tensors(device="cuda:0"):
    r = A.precompute_frequencies(64, 300)
    return r, torch.tensor(r.cos_freq, device=device), torch.tensor(r.sin_freq, device=device)


def tap(model, audio_t, label, elems):
    eng = engine(model)
    _, cos, sin = rope_tensors()
    out = torch.empty(elems, dtype=torch.float32, device="cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    rc = eng.L.a2m_debug_forward_tap(eng.h, audio_t.data_ptr(), audio_t.shape[0], cos.data_ptr(), sin.data_ptr(), 300,
                                     label.encode(), out.data_ptr(), elems, C.c_void_p(stream))
    _lib.check(eng.h, rc, "a2m_debug_forward_tap")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def debug_gemm(eng, bn, A_bf16, W_bf16, flags, bias=None, gamma=None, resid=None, want32=True, want16=False):
    M, K = A_bf16.shape
    N = W_bf16.shape[0]
    out32 = torch.zeros((M, N), dtype=torch.float32, device="cuda:0") if want32 else None
    out16 = torch.zeros((M, N), dtype=A_bf16.dtype, device="cuda:0") if want16 else None
    ptr = lambda t: None if t is None else t.data_ptr()
    stream = torch.cuda.current_stream().cuda_stream
    rc = eng.L.a2m_debug_gemm(eng.h, bn, M, N, K, A_bf16.data_ptr(), A_bf16.stride(0), W_bf16.data_ptr(), flags,
                              ptr(bias), ptr(gamma), ptr(resid), ptr(out32), ptr(out16), C.c_void_p(stream))
    _lib.check(eng.h, rc, "a2m_debug_gemm")
    torch.cuda.synchronize()
    return out32, out16
