#!/usr/bin/env python
"""npz <-> orbax checkpoint converter.  Runs where `orbax-checkpoint` is installed -- it is NOT in the
build image, so this tool is exercised there only through its npz-side helpers (tests/test_host.py).

The reference saves `eqx.filter(model_ensemble, eqx.is_inexact_array)` with an orbax CheckpointManager whose items are
('params', 'state') and whose metadata is get_model_metadata() (train.py:384-394, 799-831); it restores into a freshly built
pytree with `ocp.args.StandardRestore` (infer.py:186-207).  orbax's standard handler stores a pytree of arrays as a NESTED
DICT keyed by the pytree key path (attribute names, list indices as strings); every array has the leading ensemble axis.
`audio_to_midi_b200.infer.save_checkpoint` writes the same leaves flat, keyed by the dotted key path.  So:

    orbax -> npz :  restore the 'params' item as a raw nested dict (no target tree needed), flatten the keys with '.', save;
    npz -> orbax :  un-flatten the dotted keys into a nested dict (digit keys stay strings, as orbax writes list indices),
                    save it as the 'params' item with an empty 'state' item and the reference's metadata.

Usage:
    python tools/convert_orbax.py to-npz   <orbax_dir> <npz_dir>      # newest step
    python tools/convert_orbax.py to-orbax <npz_dir>   <orbax_dir>
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def flatten_tree(tree, prefix=""):
    """Nested dict / list of arrays -> {dotted key path: array} (the key paths of model.tree_leaves_with_path)."""
    out = {}
    if isinstance(tree, dict):
        items = tree.items()
    elif isinstance(tree, (list, tuple)):
        items = ((str(i), v) for i, v in enumerate(tree))
    else:
        if tree is not None:
            out[prefix[:-1]] = np.asarray(tree)
        return out
    for k, v in items:
        out.update(flatten_tree(v, f"{prefix}{k}."))
    return out


def unflatten_tree(leaves: dict):
    """{dotted key path: array} -> nested dict (list indices stay string keys, as orbax's standard handler names them)."""
    root: dict = {}
    for path, a in leaves.items():
        node = root
        parts = path.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = np.asarray(a)
    return root


def newest_step_dir(npz_dir: str):
    steps = sorted(int(s) for s in os.listdir(npz_dir) if s.isdigit())
    if not steps:
        raise FileNotFoundError(f"no <step>/params.npz under {npz_dir}")
    return steps[-1], os.path.join(npz_dir, str(steps[-1]))


def to_npz(orbax_dir: str, npz_dir: str):
    import orbax.checkpoint as ocp
    mgr = ocp.CheckpointManager(os.path.abspath(orbax_dir), item_names=("params", "state"))
    step = mgr.latest_step()
    if step is None:
        raise FileNotFoundError("There is no checkpoint to load!")
    restored = mgr.restore(step, args=ocp.args.Composite(params=ocp.args.StandardRestore()))
    leaves = flatten_tree(restored["params"])
    d = os.path.join(npz_dir, str(int(step)))
    os.makedirs(d, exist_ok=True)
    np.savez(os.path.join(d, "params.npz"), **{k: np.asarray(v, np.float32) for k, v in leaves.items()})
    meta = dict(mgr.metadata() or {})
    meta["ensemble_axis"] = True
    with open(os.path.join(d, "metadata.json"), "w") as f:
        json.dump(meta, f, default=str)
    print(f"step {step}: {len(leaves)} leaves -> {d}")


def to_orbax(npz_dir: str, orbax_dir: str):
    import orbax.checkpoint as ocp
    from audio_to_midi_b200.model import get_model_metadata
    step, d = newest_step_dir(npz_dir)
    with np.load(os.path.join(d, "params.npz")) as z:
        leaves = {k: z[k] for k in z.files}
    meta_file = os.path.join(d, "metadata.json")
    has_axis = True
    if os.path.exists(meta_file):
        with open(meta_file) as f:
            has_axis = bool(json.load(f).get("ensemble_axis", False))
    if not has_axis:
        leaves = {k: v[None, ...] for k, v in leaves.items()}      # the reference restores (ensemble, ...) arrays
    mgr = ocp.CheckpointManager(os.path.abspath(orbax_dir), item_names=("params", "state"), metadata=get_model_metadata())
    mgr.save(step, args=ocp.args.Composite(params=ocp.args.StandardSave(unflatten_tree(leaves)), state=ocp.args.StandardSave({})))
    mgr.wait_until_finished()
    print(f"step {step}: {len(leaves)} leaves -> {orbax_dir}")


def main(argv):
    if len(argv) != 4 or argv[1] not in ("to-npz", "to-orbax"):
        print(__doc__)
        return 2
    try:
        import orbax.checkpoint  # noqa: F401
    except ImportError:
        print("orbax-checkpoint is not installed here: run this tool on a machine that has it (SURVEY.md F1)")
        return 3
    (to_npz if argv[1] == "to-npz" else to_orbax)(argv[2], argv[3])
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
