import time, numpy as np, torch
import audio_to_midi_b200 as A
from audio_to_midi_b200 import train as T
from oracle import synth
B=64
model = A.OutputSequenceGenerator(A.model_config, key=1234)
eng = T.TrainEngine(model, 0)
cfg = T.OptimizerConfig()
rope = A.precompute_frequencies(64, 300)
host=[synth.make_windows_fast(B, 1+r) for r in range(3)]
pin_x=[torch.tensor(h).pin_memory() for h in host]
pin_y=[(torch.rand(B,250,90)*0.99).pin_memory() for _ in range(3)]
dx=[p.cuda() for p in pin_x]; dy=[p.cuda() for p in pin_y]
lr=lambda i:1e-4
for i in range(3): eng.training_step(dx[i%3],dy[i%3],rope,cfg,1e-4,dropout_rate=0.1,key=1)
torch.cuda.synchronize()
for K in (10,30):
    t0=time.perf_counter()
    for i in range(K): eng.training_step(dx[i%3],dy[i%3],rope,cfg,1e-4,dropout_rate=0.1,key=1)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print("device-resident loop K",K,dt/K*1e3,"ms/step")
    batches=[(pin_x[i%3],pin_y[i%3]) for i in range(K)]
    eng.train_pipelined(batches[:2],rope,cfg,lr,dropout_rate=0.1,key=1); torch.cuda.synchronize()
    t0=time.perf_counter(); eng.train_pipelined(batches,rope,cfg,lr,dropout_rate=0.1,key=1); torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print("pipelined K",K,dt/K*1e3,"ms/step")
    # host enqueue cost only
    t0=time.perf_counter()
    for i in range(K): eng.training_step(dx[i%3],dy[i%3],rope,cfg,1e-4,dropout_rate=0.1,key=1)
    dt=time.perf_counter()-t0; torch.cuda.synchronize()
    print("enqueue-only K",K,dt/K*1e3,"ms/step")
