#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=/root/repo:/root/repo/tests
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "forward or events or train" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-train --no-cpu --no-extra > gpurun_out/bench_nt.json 2> gpurun_out/bench_nt.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_nt.json"))
print("value",d["value"],"ms",d["ms_per_step"],"serial",d["config"]["serial_ms_per_step"])
print("e2e",{k:v for k,v in d["e2e"].items() if "value" in k})
print(d["roofline"]["families_ms"])
PY
tail -3 gpurun_out/bench_nt.err
