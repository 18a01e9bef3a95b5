#!/bin/bash
# A/B of alternative f16 builds (tools/ab_build.sh) on one GPU: short forward-only bench per variant, interleaved twice
mkdir -p gpurun_out/ab
AB=$GRAFT_REPO_ROOT/audio-to-midi_b200/_build/ab
for rep in 1 2; do
  for v in base "$@"; do
    export A2M_LIB_F16=$AB/$v.so
    python bench.py --steps 40 --warmup 5 --no-train --no-extra --no-cpu > gpurun_out/ab/${v}_$rep.json 2> gpurun_out/ab/${v}_$rep.err || tail -5 gpurun_out/ab/${v}_$rep.err
  done
done
if [ -n "$AB_CHECK" ]; then
  for v in "$@"; do
    A2M_LIB_F16=$AB/$v.so python -m pytest tests/test_gpu_forward.py -x -q -k "$AB_CHECK" 2>&1 | tail -3
  done
fi
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/ab/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        fam = d["roofline"]["families_ms"]
        print(os.path.basename(f), "value %.0f" % d["value"], "ms %.4f" % d["ms_per_step"], "serial %.4f" % d["config"]["serial_ms_per_step"], "e2e %.0f" % d["e2e"]["value"],
              " ".join("%s=%.4f" % (k.replace("_kernel", ""), v) for k, v in list(fam.items())[:8]))
    except Exception as e:
        print(f, "failed", e)
PY
