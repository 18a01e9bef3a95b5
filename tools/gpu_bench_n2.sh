#!/bin/bash
# usage: tools/gpu_bench_n2.sh N  -- both arms of bench.py on N GPUs under torch.distributed.run (as the driver launches them)
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=/root/repo
nvidia-smi topo -m 2>/dev/null | head -14
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -1 gpurun_out/bench_n$N.json | cut -c1-6000; tail -4 gpurun_out/bench_n$N.err
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
echo "ref N=$N rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -1 gpurun_out/bench_ref_n$N.json | cut -c1-800
