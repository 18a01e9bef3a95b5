"""-m gpu: the training path through the C ABI vs torch-CPU autograd of the oracle (train.py:39-62, 259-332)."""
import ctypes as C

import numpy as np
import pytest
import torch

import audio_to_midi_b200 as A
from audio_to_midi_b200 import _lib
from audio_to_midi_b200 import train as T
from gpu_util import engine, make_model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tokens,n_out,k_out,pad", [(4096, 128, 256, 0), (1000, 320, 256, 0), (2048, 512, 64, 256), (777, 64, 128, 0),
                                                     (16384, 1024, 256, 0), (3000, 256, 512, 0), (64, 128, 64, 0)])
def test_wgrad_gemm(tokens, n_out, k_out, pad):
    """dW += dY^T X on tcgen05 with both operands MN-major (gemm_wgrad.cuh) vs fp64."""
    m, _ = make_model(1, precision="bf16")      # the wgrad kernel is part of the (bf16) training path
    eng = engine(m)
    g = torch.Generator(device="cpu").manual_seed(tokens + n_out)
    dY = (torch.randn(tokens, n_out, generator=g) * 0.5).to(torch.bfloat16).cuda()
    Xfull = torch.randn(tokens, k_out + pad, generator=g).to(torch.bfloat16).cuda()
    X = Xfull[:, pad:]                       # column offset + leading dimension > k_out (the kv_down case)
    dW = torch.ones(n_out, k_out, dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    rc = eng.L.a2m_debug_wgrad(eng.h, tokens, n_out, k_out, dY.data_ptr(), dY.stride(0), X.data_ptr(), X.stride(0), dW.data_ptr(),
                               C.c_void_p(stream))
    _lib.check(eng.h, rc, "a2m_debug_wgrad")
    torch.cuda.synchronize()
    ref = 1.0 + dY.double().T @ X.double()
    err = (dW.double() - ref).abs().max().item()
    assert err < 2e-3 * np.sqrt(tokens), err


def test_backward_matches_autograd():
    """Every leaf gradient and the loss of one forward/backward of 2 windows vs torch autograd on the same weights.
    Operands are bf16 (the reference trains in fp16, train.py:36-38), accumulation fp32: 6 % relative L2 per leaf."""
    import train_util as U
    tree, audio, labels = U.setup(2)
    lref, gref, zref = U.oracle_grads(tree, audio, labels, scale=2.0)
    lcu, gcu, zcu, eng = U.cuda_grads(tree, audio, labels, scale=2.0)
    assert abs(lcu - lref) < 2e-3 * abs(lref), (lcu, lref)
    assert np.abs(zcu - zref).max() < 0.15
    rows = U.compare(gref, gcu)
    bad = [(k, round(r, 4)) for k, r, _, _ in rows if not (r < 0.06)]
    assert not bad, bad[:10]
    assert eng.launch_count() > 500


def test_backward_with_dropout_matches_autograd():
    """Dropout 0.1 on the attention weights (global and per-window local) and on the FFN output (model.py:237, 254-255):
    the oracle replays the CUDA path's counter-based masks, so losses and every leaf gradient must agree as without dropout."""
    import train_util as U
    from oracle import model_torch as MT
    tree, audio, labels = U.setup(2)
    seed = 0x1234_5678_9ABC
    masks = MT.dropout_masks(seed, 0.1, 2)
    keep = float((masks[("global", 1)] > 0).float().mean())
    assert abs(keep - 0.9) < 5e-3, keep
    lref, gref, zref = U.oracle_grads(tree, audio, labels, masks=masks)
    lcu, gcu, zcu, _ = U.cuda_grads(tree, audio, labels, dropout=0.1, seed=seed)
    l0, _, z0, _ = U.cuda_grads(tree, audio, labels)
    assert np.abs(zcu - z0).max() > 1e-2          # dropout really changes the forward
    assert abs(lcu - lref) < 2e-3 * abs(lref), (lcu, lref)
    assert np.abs(zcu - zref).max() < 0.15
    rows = U.compare(gref, gcu)
    bad = [(k, round(r, 4)) for k, r, _, _ in rows if not (r < 0.06)]
    assert not bad, bad[:10]


def test_gradient_accumulation_and_adamw():
    """Two minibatches of 1 accumulate to the gradient of the batch of 2 (train.py:283-293); AdamW + global-norm clip of
    the updates (optax.adamw then clip_by_global_norm, train.py:698-726) vs a numpy restatement on the CUDA gradients."""
    import train_util as U
    tree, audio, labels = U.setup(2)
    model, _ = make_model(7, gamma_mode="active", decoder_gain=4.0, trained_like=True)
    eng = T.TrainEngine(model, 0)
    rope = A.precompute_frequencies(64, 300)
    a, y = torch.tensor(audio).cuda(), torch.tensor(labels).cuda()
    eng.zero_grad()
    eng.forward_backward(a, y, rope)
    g_full = eng.grads.clone()
    eng.zero_grad()
    eng.forward_backward(a[:1], y[:1], rope)
    eng.forward_backward(a[1:], y[1:], rope)
    g_acc = eng.grads.clone() / 2
    rel = ((g_acc - g_full).norm() / g_full.norm()).item()
    assert rel < 2e-2, rel

    cfg = T.OptimizerConfig()
    p0 = eng.params_flat().double().cpu().numpy()
    mult = T.layer_lr_multipliers(eng.paths, cfg.layer_lr_decay)
    eng.set_lr_multipliers(mult)
    g = (eng.grads.double().cpu().numpy()) / 2.0
    eng.optimizer_step(1e-2, cfg, grad_divisor=2.0)
    torch.cuda.synchronize()
    p1 = eng.params_flat().double().cpu().numpy()
    lrm = np.ones_like(p0)
    for (o, s, m_) in zip(eng.offsets, eng.shapes, mult):
        lrm[o:o + int(np.prod(s, dtype=np.int64))] = m_
    m1 = (1 - cfg.b1) * g
    v1 = (1 - cfg.b2) * g * g
    u = -(1e-2 * lrm) * ((m1 / (1 - cfg.b1)) / (np.sqrt(v1 / (1 - cfg.b2)) + cfg.eps) + cfg.weight_decay * p0)
    norm = np.linalg.norm(u)
    u *= min(1.0, cfg.clip_norm / norm)
    assert abs(float(eng.stats[0].item()) - norm ** 2) < 1e-3 * norm ** 2
    assert float(eng.stats[1].item()) == 0.0
    assert np.abs((p1 - p0) - u).max() < 1e-6 + 1e-4 * np.abs(u).max()
    # the re-packed kernel weights follow the master parameters: a forward now differs from before the step
    eng.zero_grad()
    z2 = eng.forward_backward(a, y, rope, want_logits=True)
    torch.cuda.synchronize()
    assert torch.isfinite(z2).all()


def test_training_steps_stay_finite_and_learn():
    """compute_training_step (train.py:259-332) repeated at the bench's batch (64 windows, default init with
    gamma = 1e-6, dropout 0.1): every step's gradients are finite (the flag of train.py:320-322), repeated backward
    passes on the same batch agree (no race in the bias / gamma reductions) and the loss goes down."""
    from oracle import synth
    B = 64
    model = A.OutputSequenceGenerator(A.model_config, key=1234)
    eng = T.TrainEngine(model, 0)
    cfg = T.OptimizerConfig()
    eng.set_lr_multipliers(T.layer_lr_multipliers(eng.paths, cfg.layer_lr_decay))
    rope = A.precompute_frequencies(64, 300)
    x = torch.tensor(synth.make_windows_fast(B, 99), device="cuda:0")
    rng = np.random.Generator(np.random.PCG64(99))
    y = torch.tensor(np.clip((rng.random((B, 250, 90)) < 0.02).astype(np.float32), 0.005, 0.995), device="cuda:0")
    grads = []
    for _ in range(3):
        eng.zero_grad()
        eng.set_dropout(0.1, 5)
        eng.forward_backward(x, y, rope)
        torch.cuda.synchronize()
        assert torch.isfinite(eng.grads).all()
        grads.append(eng.grads.clone())
    for g in grads[1:]:
        assert ((g - grads[0]).norm() / grads[0].norm()).item() < 1e-3
    losses = []
    for i in range(8):
        loss, valid, _ = eng.training_step(x, y, rope, cfg, 1e-3, dropout_rate=0.1, key=3)
        assert bool(valid.item()), i
        losses.append(float(loss.item()))
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < 0.9 * losses[0], losses
    assert torch.isfinite(eng.params_flat()).all()


def test_train_pipelined_matches_step_by_step():
    """TrainEngine.train_pipelined (copy stream + double-buffered inputs + lagged loss read-back) runs the same steps as
    a plain loop of training_step on device-resident inputs: same per-step losses (the loss is summed with atomics, so to
    rounding), same parameters afterwards."""
    import train_util as U
    from oracle import params as P
    tree, audio, labels = U.setup(2)
    rope = A.precompute_frequencies(64, 300)
    cfg = T.OptimizerConfig()
    lr = lambda i: 1e-3
    batches = [(audio, labels), (audio[::-1].copy(), labels[::-1].copy()), (audio * np.float32(0.5), labels)]
    m1 = A.OutputSequenceGenerator(A.model_config, key=0).load_leaves(P.flatten(tree))
    e1 = T.TrainEngine(m1, 0)
    ref = []
    for i, (x, y) in enumerate(batches):
        loss, valid, _ = e1.training_step(torch.tensor(x).cuda(), torch.tensor(y).cuda(), rope, cfg, lr(i + 1), dropout_rate=0.1, key=5)
        ref.append(float(loss.item()))
    p1 = e1.params_flat().cpu().numpy()
    del e1
    m2 = A.OutputSequenceGenerator(A.model_config, key=0).load_leaves(P.flatten(tree))
    e2 = T.TrainEngine(m2, 0)
    pinned = [(torch.tensor(x).pin_memory(), torch.tensor(y).pin_memory()) for x, y in batches]
    got = e2.train_pipelined(pinned, rope, cfg, lr, first_step=1, dropout_rate=0.1, key=5)
    p2 = e2.params_flat().cpu().numpy()
    assert len(got) == 3
    # gradients are reduced with atomics, and Adam's normalised update amplifies their rounding: a few 1e-4 per step
    assert np.allclose(got, ref, rtol=2e-3), (got, ref)
    assert np.abs(p1 - p2).max() < 5e-3 * max(1.0, np.abs(p1).max())


def test_gradient_bucket_is_final_when_its_event_fires():
    """SURVEY 8e (all-reduce bucketed and overlapped with the backward): bucket 0 of the gradient blob -- final norm,
    transformer, decoder, the tail of the leaf order -- is scattered by a step INSIDE the backward graph and announced by an
    event-record node.  A second stream that waits on it must see exactly the final values of that range while the CNN
    backward is still running; the buckets tile the blob."""
    import train_util as U
    tree, audio, labels = U.setup(4)
    m = A.OutputSequenceGenerator(A.model_config, key=0).load_leaves(__import__("oracle.params", fromlist=["flatten"]).flatten(tree))
    eng = T.TrainEngine(m, 0)
    rope = A.precompute_frequencies(64, 300)
    buckets = eng.grad_buckets()
    assert len(buckets) == 2 and buckets[1][0] == 0 and buckets[1][1] == buckets[0][0] and buckets[0][1] == eng.n_params
    first_tail = min(o for p, o in zip(eng.paths, eng.offsets) if not p.startswith("layers."))
    assert buckets[0][0] == first_tail
    assert (buckets[0][1] - buckets[0][0]) > 0.75 * eng.n_params          # most of the bytes can overlap
    x, y = torch.tensor(audio).cuda(), torch.tensor(labels).cuda()
    side = torch.cuda.Stream()
    for _ in range(3):                                                      # first call captures the graph, later ones replay it
        eng.zero_grad()
        eng.forward_backward(x, y, rope)
        _lib.check(eng.h, eng.L.a2m_stream_wait_grad_bucket(eng.h, 0, C.c_void_p(side.cuda_stream)), "wait bucket")
        with torch.cuda.stream(side):
            snap = eng.grads[buckets[0][0]:].clone()
            head_early = eng.grads[:buckets[0][0]].clone()
        torch.cuda.synchronize()
        final = eng.grads[buckets[0][0]:]
        assert torch.equal(snap, final)
        assert float(final.abs().max()) > 0
    # the CNN bucket was still being produced when bucket 0 fired (otherwise nothing overlaps)
    assert not torch.equal(head_early, eng.grads[:buckets[0][0]])


def test_train_then_predict_then_train_and_sync():
    """the reference loop validates every N steps (train.py:396-437).  Inference through the model while a
    TrainEngine is alive must (a) run on the weights being trained, (b) leave the training session intact, and
    sync_to_model / save_checkpoint must store the TRAINED weights."""
    import train_util as U
    from oracle import params as P
    tree, audio, labels = U.setup(2)
    rope = A.precompute_frequencies(64, 300)
    cfg = T.OptimizerConfig()
    model = A.OutputSequenceGenerator(A.model_config, key=0).load_leaves(P.flatten(tree))
    x, y = torch.tensor(audio).cuda(), torch.tensor(labels).cuda()
    _, p_before = model.predict(None, x, rope)                       # the model's own inference handle
    p_before = p_before.clone()
    eng = T.TrainEngine(model, 0)
    _, p_live0 = model.predict(None, x, rope)                        # now the trainer's handle, same weights (un-folded kv path)
    assert (p_live0 - p_before).abs().max().item() < 3e-2    # the model's own handle has f16 operands, the trainer's bf16
    for i in range(3):
        loss, valid, _ = eng.training_step(x, y, rope, cfg, 1e-2, key=1)
        assert bool(valid.item())
    _, p_live = model.predict(None, x, rope)                         # validation inside the loop: sees the trained weights
    assert (p_live - p_before).abs().max().item() > 1e-3
    loss2, valid2, _ = eng.training_step(x, y, rope, cfg, 1e-2, key=1)    # ... and the session survives it
    assert bool(valid2.item()) and np.isfinite(float(loss2.item())) and float(loss2.item()) < float(loss.item()) * 1.5
    # a second model can be resident next to the trainer without disturbing either
    other, _ = make_model(3)
    _, p_other = other.predict(None, x, rope)
    assert torch.isfinite(p_other).all()
    loss3, valid3, _ = eng.training_step(x, y, rope, cfg, 1e-2, key=1)
    assert bool(valid3.item())
    trained = eng.params_tree()
    eng.sync_to_model()
    leaves = dict(model.tree_leaves_with_path())
    assert all(np.array_equal(np.asarray(leaves[k]), trained[k]) for k in trained)
    assert not np.array_equal(np.asarray(leaves["decoder.decoder_pooling.weight"]), P.flatten(tree)["decoder.decoder_pooling.weight"])
    _, p_sync = model.predict(None, x, rope)                         # still the trainer's handle
    eng.close()
    _, p_closed = model.predict(None, x, rope)                       # the model's own handle, re-loaded with the trained leaves
    assert (p_closed - p_sync).abs().max().item() < 3e-2


def test_non_finite_gradients_leave_state_untouched():
    """ADVICE r1 / train.py:369-377: with an inf or NaN in the gradients the optimizer step is a device-side no-op -- master
    weights and both moments are bit-identical afterwards, stats[1] counts the bad entries; snapshot()/restore() roll a
    session back; train_pipelined reports the invalid step, halves the loss scale and takes the update count back."""
    import train_util as U
    from oracle import params as P
    tree, audio, labels = U.setup(2)
    rope = A.precompute_frequencies(64, 300)
    cfg = T.OptimizerConfig()
    model = A.OutputSequenceGenerator(A.model_config, key=0).load_leaves(P.flatten(tree))
    eng = T.TrainEngine(model, 0)
    x, y = torch.tensor(audio).cuda(), torch.tensor(labels).cuda()
    eng.training_step(x, y, rope, cfg, 1e-3, key=1)
    snap = eng.snapshot()
    p0, m0, v0, step0 = snap
    eng.zero_grad()
    eng.forward_backward(x, y, rope)
    eng.grads[12345] = float("inf")
    eng.grads[-7] = float("nan")
    eng.optimizer_step(1e-3, cfg)
    torch.cuda.synchronize()
    assert float(eng.stats[1].item()) == 2.0
    p1, m1, v1, _ = eng.snapshot()
    assert torch.equal(p0, p1) and torch.equal(m0, m1) and torch.equal(v0, v1)
    eng.step_count -= 1
    loss, valid, _ = eng.training_step(x, y, rope, cfg, 1e-3, key=1)          # a clean step moves everything again
    assert bool(valid.item())
    p2, m2, v2, _ = eng.snapshot()
    assert not torch.equal(p0, p2) and not torch.equal(m0, m2)
    eng.restore(snap)
    p3, m3, v3, step3 = eng.snapshot()
    assert torch.equal(p0, p3) and torch.equal(m0, m3) and torch.equal(v0, v3) and step3 == step0
    # host-fed loop: the second batch carries an inf sample -> its step is reported invalid and skipped, the scale halves
    bad = audio.copy()
    bad[0, 0, 100] = np.inf
    batches = [(torch.tensor(a_).pin_memory(), torch.tensor(labels).pin_memory()) for a_ in (audio, bad, audio)]
    before = eng.step_count
    losses, valids, scale = eng.train_pipelined(batches, rope, cfg, lambda i: 1e-3, first_step=before, grad_scale=4.0, return_valid=True)
    assert valids == [True, False, True] and scale == 2.0 and eng.step_count == before + 2
    assert np.isfinite(losses[0]) and np.isfinite(losses[2])
    assert torch.isfinite(eng.params_flat()).all()


def test_model_call_with_dropout_and_arbitrary_cotangent():
    """b1: the reference's training call shape `vmap(model, (0, None, None, 0, None))(audio, state, rope, keys, True)`
    (train.py:56-58) runs a2m_forward_train with dropout masks seeded by the keys; a2m_backward_dlogits with the BCE cotangent
    reproduces a2m_backward (the custom_vjp pair a JAX host binds)."""
    import train_util as U
    from audio_to_midi_b200.model import fold_key
    from oracle import model_torch as MT
    from oracle import params as P
    tree, audio, labels = U.setup(2)
    rope = A.precompute_frequencies(64, 300)
    model = A.OutputSequenceGenerator(A.model_config, key=0).load_leaves(P.flatten(tree))
    x, y = torch.tensor(audio).cuda(), torch.tensor(labels).cuda()
    keys = np.array([[0, 11], [0, 12]], np.uint32)
    call = A.vmap(model, in_axes=(0, None, None, 0, None), out_axes=(0, None), axis_name="batch")
    (logits, probs), state = call(x, None, rope, keys, True)
    assert state is None and tuple(logits.shape) == (2, 250, 90)
    masks = MT.dropout_masks(fold_key(keys), 0.1, 2)
    with torch.no_grad():
        zref, pref = MT.forward(MT.to_torch(tree), torch.tensor(audio), masks=masks)
    assert (logits.cpu() - zref).abs().max().item() < 0.15 and (probs.cpu() - pref).abs().max().item() < 3e-2
    (z_eval, _), _ = model(x, None, rope)                              # inference call: no dropout, different logits
    assert (z_eval - logits).abs().max().item() > 1e-3
    eng = model._live_trainer(0)
    assert eng is not None
    eng.zero_grad()
    eng.backward(y, scale=2.0)
    g_bce = eng.grads.clone()
    (logits2, _), _ = call(x, None, rope, keys, True)                  # same keys -> same masks -> same forward
    assert torch.equal(logits2, logits)
    dz = (torch.sigmoid(logits2) - y) * (2.0 / 2)                      # d/dz of mean_b(sum BCE * scale), train.py:43-47,61-62
    eng.zero_grad()
    eng.backward_dlogits(dz)
    rel = ((eng.grads - g_bce).norm() / g_bce.norm()).item()
    assert rel < 2e-2, rel
