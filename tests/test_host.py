"""CPU-side tests of the product's host logic: the C ABI loads and exports what include/a2m.h declares, the
pytree mirrors the reference layout, modelutil (C++) matches the oracle bit for bit, window sharding, and the
product refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import audio_to_midi_b200 as A
from audio_to_midi_b200 import _lib
from oracle import events as E
from oracle import params as P

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_abi_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "a2m.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(a2m_\w+|extract_midi_events|free_midi_events)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} not exported"


def test_pytree_matches_reference_layout():
    m = A.OutputSequenceGenerator(A.model_config, key=1)
    leaves = m.tree_leaves_with_path()
    ref = P.flatten(P.init_params(1))
    assert [p for p, _ in leaves] == list(ref.keys())
    assert all(np.shape(a) == np.shape(ref[p]) for p, a in leaves)
    assert sum(np.asarray(a).size for _, a in leaves) == 11_606_269
    top = [n for n in m._fields]
    assert top == ["layers", "norm", "transformer_projection", "transformer", "decoder"]      # model.py:673-678
    with pytest.raises(KeyError):
        m.load_leaves({"norm.weight": np.ones(256, np.float32)})
    bad = dict(ref)
    bad["norm.weight"] = np.ones(255, np.float32)
    with pytest.raises(ValueError):
        m.load_leaves(bad)


def test_rope_table_matches_oracle():
    from oracle import model_np as M
    r = A.precompute_frequencies(64, 300)
    cos, sin = M.precompute_frequencies(64, 300, dtype=np.float32)
    assert r.cos_freq.shape == (300, 32) and r.cos_freq.dtype == np.float32
    assert np.array_equal(r.cos_freq, cos) and np.array_equal(r.sin_freq, sin)


def test_modelutil_bit_exact_vs_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "events.npz"))
    for name, ov in (("ov050", 0.5), ("ov025", 0.25), ("ov000", 0.0)):
        st = A.modelutil.stitch_probs(g["probs"], ov, 0.02)
        assert np.array_equal(st, g["stitched_" + name], equal_nan=True)
        ev = A.modelutil.extract_events(st)
        assert ev == [tuple(r) for r in g["events_" + name].tolist()]
        fr = A.modelutil.to_frame_events([ev], st.shape[0])[0]
        assert np.array_equal(fr, g["frames_" + name])


def test_modelutil_edge_cases():
    assert A.modelutil.extract_events(np.zeros((0, 90), np.float32)) == []
    assert A.modelutil.extract_events(np.zeros((5, 90), np.float32)) == []
    one = np.zeros((1, 90), np.float32)
    one[0, 3] = 0.9
    assert A.modelutil.extract_events(one) == E.extract_events(one) == [(0, 3, 1, 7)]
    rng = np.random.Generator(np.random.PCG64(3))
    for frames in (2, 7, 13, 64):
        p = rng.uniform(size=(frames, 90)).astype(np.float32)
        assert A.modelutil.extract_events(p) == E.extract_events(p)
    w1 = rng.uniform(size=(1, 250, 90)).astype(np.float32)
    assert np.array_equal(A.modelutil.stitch_probs(w1, 0.5, 0.02), w1[0])
    assert A.modelutil.to_frame_events([[]], 10)[0].sum() == 0
    with pytest.raises(ValueError):
        A.modelutil.to_frame_events([[(0, 90, 5, 7)]], 10)


def test_ios_c_abi_f16_strided():
    """extract_midi_events (cbinds.rs:51-91): f16 data, strides in elements, stitch then extract."""
    L = _lib.lib()
    rng = np.random.Generator(np.random.PCG64(8))
    probs = rng.uniform(size=(3, 250, 90)).astype(np.float16)
    padded = np.zeros((3, 250, 96), np.float16)          # non-contiguous rows: stride 96 elements
    padded[:, :, :90] = probs
    w = _lib.MLMultiArrayWrapper3()
    w.strides[:] = [250 * 96, 96, 1]
    w.dims[:] = [3, 250, 90]
    w.data = padded.ctypes.data
    lst = L.extract_midi_events(w, 0.5, 0.02)
    got = [(lst.contents.ptr[i].attack_time, lst.contents.ptr[i].note, lst.contents.ptr[i].duration,
            lst.contents.ptr[i].velocity) for i in range(lst.contents.length)]
    L.free_midi_events(lst)
    L.free_midi_events(None)                              # null is tolerated (cbinds.rs:82)
    want = E.extract_events(E.stitch_probs(probs.astype(np.float32), 0.5, 0.02))
    assert got == want
    assert C.sizeof(_lib.MidiEvent) == 32                 # #[repr(C)] {u64, u8, u64, u8}


def test_slice_and_shard():
    clip = np.zeros((2, 9_600_000), np.float32)
    w, dur = A.slice_windows(clip, overlap=0.5)
    assert w.shape == (134, 2, 80000) and dur == 5.0
    ref = E.slice_windows(np.arange(2 * 200_000, dtype=np.float32).reshape(2, -1), overlap=0.25)
    got, _ = A.slice_windows(np.arange(2 * 200_000, dtype=np.float32).reshape(2, -1), overlap=0.25)
    assert np.array_equal(ref, got)
    for n, ws in ((134, 8), (127, 8), (5, 8), (64, 4), (1, 2)):
        blocks = [A.shard_windows(n, ws, r) for r in range(ws)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_no_cpu_fallback():
    """Without an sm_100 device the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = A.OutputSequenceGenerator(A.model_config, key=1)
    with pytest.raises(_lib.A2mError):
        m.predict(None, np.zeros((1, 2, 80000), np.float32), A.precompute_frequencies(64, 300))
    with pytest.raises(_lib.A2mError):     # the training-mode call shape (train.py:56-58) exists and is GPU-only as well
        m(np.zeros((2, 80000), np.float32), None, A.precompute_frequencies(64, 300), key=1, enable_dropout=True)
    with pytest.raises(ValueError):
        m.predict(None, np.zeros((2, 1000), np.float32), A.precompute_frequencies(64, 300))


def _shard_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 37
    a, b = A.shard_windows(n, world, rank)
    mine = torch.zeros(n)
    mine[a:b] = torch.arange(a, b, dtype=torch.float32) + 1      # stand-in for per-window results
    dist.all_reduce(mine)                                        # gather-by-sum: every window owned exactly once
    q.put((rank, mine.tolist()))
    dist.destroy_process_group()


def test_window_sharding_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for _, vals in res:
        assert vals == [float(i + 1) for i in range(37)]


def _allreduce_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audio_to_midi_b200 import train as T
    lo, hi = T.shard_batch(8, world, rank)
    g = torch.full((5,), float(rank + 1))
    loss = torch.tensor([float(hi)])
    T.allreduce_mean_([g, loss])
    q.put((rank, lo, hi, g.tolist(), loss.item()))
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    """Host logic of the data-parallel training step (batch split + gradient mean) on 2 CPU ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    ps = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert [(g[1], g[2]) for g in got] == [(0, 4), (4, 8)]
    for g in got:
        assert g[3] == [1.5] * 5 and g[4] == 6.0


def test_lr_schedule_and_multipliers():
    """create_learning_rate_schedule (train.py:454-466) and the layer-wise decay labels of setup_optimizers (train.py:648-704)."""
    from audio_to_midi_b200 import train as T
    sch = T.create_learning_rate_schedule(1e-4, 1000, 200_000)
    assert sch(0) == 0.0 and abs(sch(500) - 5e-5) < 1e-12 and abs(sch(1000) - 1e-4) < 1e-12
    assert abs(sch(1000 + 100_000) - 5e-5) < 1e-9 and sch(1000 + 200_000) < 1e-12
    paths = ["layers.0.layers.0.conv.weight", "layers.0.layers.3.gamma", "layers.6.layers.3.gamma", "norm.weight",
             "transformer.layers.local_attention.attention_norm.weight"]
    m = T.layer_lr_multipliers(paths, 0.7)
    assert m[-1] == 1.0 and m[3] == 1.0 and m[2] == 1.0          # deepest conv layer: decay ** 0
    assert abs(m[0] - 0.7 ** 39) < 1e-9 and abs(m[1] - 0.7 ** 36) < 1e-9


def test_midi_writer_round_trip(tmp_path):
    """write_midi_file (infer.py:46-83): header, tempo / time-signature metas, tick conversion of mido.second2tick,
    velocity scaling, note_off before note_on at equal ticks, end_of_track."""
    from audio_to_midi_b200 import infer
    events = [(0, 39, 25, 7), (25, 39, 10, 7), (10, 60, 5, 10), (100, 0, 1, 1)]
    f = tmp_path / "out.mid"
    infer.write_midi_file(events, 0.02, str(f))
    raw = f.read_bytes()
    assert raw[:14] == b"MThd" + (6).to_bytes(4, "big") + b"\x00\x01\x00\x01\x01\xe0"
    assert raw[22:29] == b"\x00\xff\x51\x03\x07\xa1\x20"            # set_tempo 500000 us
    assert raw[29:37] == b"\x00\xff\x58\x04\x04\x02\x18\x08"        # 4/4, 24 clocks per click, 8 32nds per beat
    assert raw[-4:] == b"\x00\xff\x2f\x00"
    notes = infer.read_midi_notes(str(f))
    tick = lambda frame: int(round(frame * 0.02 / (500000e-6 / 480)))
    want = sorted([(tick(a), "note_on", k + 21, int(round(v / 10 * 127))) for a, k, d, v in events] +
                  [(tick(a + d), "note_off", k + 21, int(round(v / 10 * 127))) for a, k, d, v in events])
    assert notes == want
    i_off = notes.index((tick(25), "note_off", 60, 89))
    assert notes[i_off + 1] == (tick(25), "note_on", 60, 89)           # release of the first note precedes the re-attack
    assert tick(25) == 480                                              # 0.5 s at 120 bpm = one beat


def test_checkpoint_npz_round_trip(tmp_path):
    """save_checkpoint / load_newest_checkpoint: every pytree leaf under its key path, newest step wins."""
    from audio_to_midi_b200 import infer
    m0 = A.OutputSequenceGenerator(A.model_config, key=3)
    m1 = A.OutputSequenceGenerator(A.model_config, key=4)
    infer.save_checkpoint(m0, str(tmp_path), 100)
    infer.save_checkpoint(m1, str(tmp_path), 2000)
    got, state = infer.load_newest_checkpoint(str(tmp_path))
    assert state is None
    a, b = dict(m1.tree_leaves_with_path()), dict(got.tree_leaves_with_path())
    assert list(a) == list(b) and len(a) == 450
    assert all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for k in a)
    assert not np.array_equal(np.asarray(dict(m0.tree_leaves_with_path())["decoder.decoder_pooling.weight"]),
                              np.asarray(b["decoder.decoder_pooling.weight"]))


def test_checkpoint_ensemble_axis_and_converter_helpers(tmp_path):
    """Checkpointed arrays carry the reference's leading ensemble axis (train.py:788-795: leaves are (1, ...), transformer
    leaves (1, 8, ...)); load_newest_checkpoint selects a member (infer.py:213-221) and also accepts axis-free files.  The
    orbax converter's npz-side helpers (nested dict <-> dotted key paths) round-trip the pytree."""
    import importlib.util
    from audio_to_midi_b200 import infer
    m = A.OutputSequenceGenerator(A.model_config, key=5)
    d = infer.save_checkpoint(m, str(tmp_path / "a"), 7)
    with np.load(os.path.join(d, "params.npz")) as z:
        assert z["norm.weight"].shape == (1, 256)
        assert z["transformer.layers.local_attention.attention_block.self_attention.kv_down_proj.weight"].shape == (1, 8, 64, 256)
        stacked = {k: np.concatenate([z[k], z[k] + 1.0]) for k in z.files}          # a two-member ensemble
    got, _ = infer.load_newest_checkpoint(str(tmp_path / "a"))
    a, b = dict(m.tree_leaves_with_path()), dict(got.tree_leaves_with_path())
    assert all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for k in a)
    os.makedirs(tmp_path / "b" / "9")
    np.savez(tmp_path / "b" / "9" / "params.npz", **stacked)
    got1, _ = infer.load_newest_checkpoint(str(tmp_path / "b"), ensemble_size=2, ensemble_select=1)
    b1 = dict(got1.tree_leaves_with_path())
    assert all(np.array_equal(np.asarray(a[k]) + 1.0, np.asarray(b1[k])) for k in a)
    with pytest.raises(IndexError):
        infer.load_newest_checkpoint(str(tmp_path / "b"), ensemble_select=2)
    infer.save_checkpoint(m, str(tmp_path / "c"), 3, ensemble_axis=False)           # round-1 layout still loads
    got2, _ = infer.load_newest_checkpoint(str(tmp_path / "c"))
    assert all(np.array_equal(np.asarray(a[k]), np.asarray(v)) for k, v in got2.tree_leaves_with_path())
    spec = importlib.util.spec_from_file_location("convert_orbax", os.path.join(ROOT, "tools", "convert_orbax.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    nested = conv.unflatten_tree({k: np.asarray(v) for k, v in a.items()})
    assert set(nested) == {"layers", "norm", "transformer", "decoder"}                 # model.py:673-678 (transformer_projection is None)
    assert set(nested["layers"]["5"]["layers"]) == {str(i) for i in range(22)}
    flat = conv.flatten_tree(nested)
    assert set(flat) == set(a) and all(np.array_equal(flat[k], np.asarray(a[k])) for k in a)


def test_config_struct_and_key_folding():
    """A2mConfig mirrors model_config (model.py:20-34); PRNG keys fold into one dropout seed deterministically."""
    from audio_to_midi_b200 import model as Mo
    c = Mo.default_config_struct(3)
    assert c.device == 3 and list(c.dims)[:7] == [4, 8, 16, 32, 64, 128, 256] and list(c.depths)[:7] == [3, 3, 3, 3, 3, 21, 3]
    assert (c.num_transformer_layers, c.num_transformer_heads, c.attention_size, c.compressed_attention_kv_size,
            c.transformer_intermediate, c.cnn_hidden_expansion_x2) == (8, 4, 64, 64, 512, 4)
    assert Mo.fold_key(None) == 0 and Mo.fold_key(5) == Mo.fold_key(np.array([5])) != Mo.fold_key(6)
    assert Mo.fold_key(np.array([[0, 1], [0, 2]], np.uint32)) != Mo.fold_key(np.array([[0, 2], [0, 1]], np.uint32))
    assert 0 <= Mo.fold_key(-3) < 2 ** 64


def test_lr_schedule_is_zero_based():
    """optax.scale_by_schedule evaluates schedule(count) BEFORE incrementing: the first update of the warm-up has lr 0."""
    from audio_to_midi_b200 import train as T
    s = T.create_learning_rate_schedule(1e-4, 1000, 200_000)
    assert s(0) == 0.0 and abs(s(1) - 1e-7) < 1e-15 and abs(s(1000) - 1e-4) < 1e-12
    assert s(1001) < 1e-4 and abs(s(1000 + 200_000)) < 1e-12


def _gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audio_to_midi_b200 import infer as I
    n = 7                                                        # ragged: blocks of 3, 2, 2 over three ranks
    lo, hi = I.shard_windows(n, world, rank)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None, None] * torch.ones(1, 4, 3)     # window w carries the value w
    full = I.gather_window_blocks(local, n, world, rank)
    q.put((rank, lo, hi, full[:, 0, 0].tolist(), tuple(full.shape)))
    dist.destroy_process_group()


def test_rank_ordered_gather_world3_gloo():
    """The gather that follows the sharded forward of configs 3 and 5 (infer.gather_window_blocks): ragged contiguous blocks,
    padded to the largest, one equal-size all_gather, concatenated in rank order == window order, on every rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 150
    ps = [ctx.Process(target=_gather_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert [(g[1], g[2]) for g in got] == [(0, 3), (3, 5), (5, 7)]
    for g in got:
        assert g[3] == [float(i) for i in range(7)] and g[4] == (7, 4, 3)


def test_xla_ffi_source_type_checks():
    """csrc/a2m_xla_ffi.cc (the jax.ffi handlers) cannot be built here -- jaxlib's headers are absent (F1) -- but it is
    type-checked against a stub of the part of xla/ffi/api/ffi.h it uses (tests/stubs): every handler Impl must be callable with
    exactly the argument / result / attribute types its binding declares, and every a2m_* call must match include/a2m.h."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    src = os.path.join(ROOT, "audio-to-midi_b200", "csrc", "a2m_xla_ffi.cc")
    res = subprocess.run([gxx, "-fsyntax-only", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "tests", "stubs"),
                          "-I/usr/local/cuda/include", src], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    text = open(src).read()
    for sym in ("A2mForward", "A2mForwardTrain", "A2mBackward", "A2mLossAndGrad", "A2mAllReduce", "A2mAdamW"):
        assert f"XLA_FFI_DEFINE_HANDLER_SYMBOL({sym}," in text
    binding = open(os.path.join(ROOT, "audio-to-midi_b200", "jax_binding.py")).read()
    for sym in ("A2mForward", "A2mForwardTrain", "A2mBackward", "A2mLossAndGrad", "A2mAllReduce", "A2mAdamW"):
        assert f'"{sym}"' in binding


def test_operand_format_variants_and_host_rounding():
    """The two builds of the library (ptx.cuh: A2M_OP_F16): same exports, they name their tensor-core operand format, and the
    host-side weight rounding of the f16 variant is IEEE round-to-nearest-even binary16 (numpy's), of the bf16 variant the
    top 16 bits with RNE -- including subnormals, overflow to infinity, signed zero and NaN."""
    import ctypes as C
    rng = np.random.Generator(np.random.PCG64(3))
    x = np.concatenate([rng.normal(0, 1, 4000), rng.normal(0, 1e-6, 2000), rng.normal(0, 3e4, 2000),
                        np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, -70000.0, 2.0 ** -24, 2.0 ** -25, 1.5 * 2.0 ** -25, 2.0 ** -14,
                                  np.inf, -np.inf, np.nan, 1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11])]).astype(np.float32)
    for precision in ("bf16", "f16"):
        L = _lib.lib(precision)
        assert L.a2m_operand_format().decode() == precision
        for name in _lib.EXPORTS:
            assert hasattr(L, name), (precision, name)
        out = np.zeros(x.size, np.uint16)
        assert L.a2m_debug_round_operand(x.ctypes.data, out.ctypes.data, x.size) == 0
        if precision == "f16":
            with np.errstate(over="ignore"):
                ref = x.astype(np.float16)
            got = out.view(np.float16)
            nan = np.isnan(ref)
            assert np.array_equal(np.isnan(got), nan)
            assert np.array_equal(got[~nan].view(np.uint16), ref[~nan].view(np.uint16))
        else:
            import torch
            ref = torch.tensor(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
            nan = np.isnan(x)
            assert np.array_equal(out[~nan], ref[~nan])
            assert np.all((out[nan] & 0x7F80) == 0x7F80) and np.all(out[nan] & 0x7F)


def test_change_fp_precision_selects_the_operand_variant():
    """infer.py:27-32: change_fp_precision(model, dtype) -> the operand format of the model's inference handles."""
    m = A.OutputSequenceGenerator(A.model_config, key=1)
    assert m.precision in ("bf16", "f16")
    assert A.change_fp_precision(m, np.float16) is m and m.precision == "f16"
    assert A.change_fp_precision(m, "bfloat16").precision == "bf16"
    assert A.change_fp_precision(m, np.float32).precision == "f16"
    with pytest.raises(ValueError):
        A.change_fp_precision(m, np.int8)


def test_balanced_batches_cover_the_range_evenly():
    """infer.balanced_batches: the cut of a rank's windows for the two compute lanes -- contiguous, complete, at most max_batch
    per batch, an even number of batches when more than one, sizes equal up to one window."""
    from audio_to_midi_b200.infer import balanced_batches
    assert balanced_batches(0, 134, 72) == [(0, 67), (67, 134)]
    assert balanced_batches(0, 134, 64) == [(0, 33), (33, 67), (67, 100), (100, 134)]
    assert balanced_batches(5, 5, 64) == [] and balanced_batches(3, 4, 64) == [(3, 4)]
    for lo, n, mb in [(0, 1, 1), (7, 64, 64), (0, 65, 64), (10, 512, 64), (0, 3, 1), (0, 200, 72), (0, 17, 72), (2, 129, 64)]:
        spans = balanced_batches(lo, lo + n, mb)
        assert spans[0][0] == lo and spans[-1][1] == lo + n
        assert all(b == c for (_a, b), (c, _d) in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) <= mb and max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
        assert len(spans) == 1 or len(spans) % 2 == 0 or len(spans) == n


def test_empty_batch_maps_to_empty_outputs_without_a_device():
    """jax.vmap(model.predict) over a zero-length batch returns zero-length outputs (infer.py:40); the mirror does the same before
    it ever asks for a GPU, so the edge case is covered on the CPU runner too."""
    import audio_to_midi_b200 as A
    model = A.OutputSequenceGenerator(A.model_config, key=3)
    rope = A.precompute_frequencies(64, 300)
    logits, probs = model.predict(None, np.zeros((0, 2, 80000), np.float32), rope)
    assert logits.shape == probs.shape == (0, 250, 90) and logits.dtype == np.float32
    with pytest.raises(ValueError):
        model.predict(None, np.zeros((0, 2, 1000), np.float32), rope)
