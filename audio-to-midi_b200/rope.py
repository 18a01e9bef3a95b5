"""RoPE frequency table: host mirror of the reference's rope.py (RopeFreqs rope.py:5-9,
precompute_frequencies rope.py:12-22).  The rotation itself (calculate_rope, rope.py:25-53) runs inside
the CUDA kernels (GEMM epilogue for global attention, attn_local_kernel for local windows)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class RopeFreqs:
    cos_freq: np.ndarray  # (max_pos, dim // 2) fp32
    sin_freq: np.ndarray


def precompute_frequencies(dim: int, max_pos: int, theta: float = 10000.0) -> RopeFreqs:
    """Same arithmetic as the reference, in fp32."""
    f32 = np.float32
    inv_freq = f32(1.0) / (f32(theta) ** (np.arange(0, dim, 2, dtype=f32)[: dim // 2] / f32(dim)))
    t = np.arange(0, max_pos, dtype=f32)
    freqs = np.outer(t, inv_freq).astype(f32)
    return RopeFreqs(cos_freq=np.cos(freqs).astype(f32), sin_freq=np.sin(freqs).astype(f32))
