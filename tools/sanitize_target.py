"""Small end-to-end exercise of every entry point, for compute-sanitizer (memcheck): python tools/sanitize_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_to_midi_b200 as A
from audio_to_midi_b200 import infer as I, train as T
from oracle import synth

rope = A.precompute_frequencies(64, 300)
audio, labels = synth.make_windows(3, 5, with_labels=True)
x = torch.tensor(audio).cuda()
for precision in ("f16", "bf16"):
    m = A.OutputSequenceGenerator(A.model_config, key=3)
    m.precision = precision
    lg, pr = m.predict(None, x, rope)
    outs = m.predict_many(None, [x[:2], x[2:], x[:1]], rope)
    host = list(m.predict_pipelined([audio.astype(np.float16), audio[:2]], rope, copy=True, want_logits=False, probs_dtype=np.float16))
    torch.cuda.synchronize()
    print(precision, float(pr.mean()), float(outs[1][1].mean()), host[0][1].shape)
m = A.OutputSequenceGenerator(A.model_config, key=3)
met = I.detailed_event_loss_device(m, pr, torch.tensor(labels).cuda())
raw = np.asarray(synth.make_clip(17.0, 3), np.float32)
ev, st, probs = I.transcribe_clip(m, raw, overlap=0.5, max_batch=2)
print("clip", len(ev), st.shape, probs.shape, met.cpu().numpy()[0])
eng = T.TrainEngine(m, 0)
cfg = T.OptimizerConfig()
y = torch.tensor(labels).cuda()
for i in range(2):
    loss, valid, _ = eng.training_step(x[:2], y[:2], rope, cfg, 1e-3, dropout_rate=0.1, key=1)
torch.cuda.synchronize()
print("train", float(loss.item()), bool(valid.item()))
_, p2 = m.predict(None, x, rope)
eng.zero_grad(); eng.forward_train(x[:2], rope); eng.backward_dlogits(torch.zeros(2, 250, 90, device="cuda") + 1e-3)
torch.cuda.synchronize()
print("done", float(p2.mean()), float(eng.grads.abs().max()))
