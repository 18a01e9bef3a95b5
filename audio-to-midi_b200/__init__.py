"""B200-native implementation of the audio-to-midi hot path (batched model forward).

Layout:  csrc/ (CUDA kernels + C ABI, include/a2m.h)  model.py / rope.py / modelutil.py / infer.py
(host mirrors of the reference modules with the same names).  Import as ``audio_to_midi_b200``.
"""
from . import modelutil  # noqa: F401
from .infer import predict_and_stitch, shard_windows, slice_windows  # noqa: F401
from .model import OutputSequenceGenerator, change_fp_precision, get_model_metadata, model_config, pinned_empty, vmap  # noqa: F401
from .rope import RopeFreqs, precompute_frequencies  # noqa: F401
