"""CPU oracle for the audio-to-midi hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's batched
model forward (``/root/reference/model.py``, ``rope.py``), its training loss
(``train.py:39-62``) and the post-processing that consumes its output
(``rust-plugins/src/common.rs``, ``python.rs:423-447``).

PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors for
this path and neither JAX/equinox nor a Rust toolchain exist in
this image, so the reference itself cannot be executed here.  The library
semantics this restatement assumes are listed in SURVEY.md §8(c); the oracle is
pinned only against itself (fp64 numpy vs. an independent fp32 torch twin) and
against closed-form self-checks (tests/test_oracle_*.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the
product path (``audio-to-midi_b200/``) never does.
"""
