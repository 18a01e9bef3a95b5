"""Fills the round-2 results table of DESIGN.md (marker R02_TABLE, or the previous table between the markers) from profiles/r02_bench*.json."""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def load(n):
    f = os.path.join(ROOT, "profiles", "r02_bench%s.json" % ("" if n == 1 else "_n%d" % n))
    return json.loads(open(f).read().strip().splitlines()[-1])
d = {n: load(n) for n in (1, 2, 4, 8)}
ref = {n: json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_ref%s.json" % ("" if n == 1 else "_n%d" % n))).read().strip().splitlines()[-1]) for n in (1, 2, 4, 8)}
k = lambda v: f"{v / 1e3:.1f} k" if v < 1e6 else f"{v / 1e6:.2f} M"
rows = []
rows.append("| | 1 GPU | 2 GPUs | 4 GPUs | 8 GPUs |")
rows.append("|---|---|---|---|---|")
rows.append("| forward `value`, audio-s/s (two lanes) | " + " | ".join(f"**{k(d[n]['value'])}** ({d[n]['ms_per_step']:.3f} ms/step)" for n in d) + " |")
rows.append("| … one step after the other (`serial_value`) | " + " | ".join(f"{k(d[n]['config']['serial_value'])} ({d[n]['config']['serial_ms_per_step']:.3f} ms)" for n in d) + " |")
rows.append("| forward `e2e`, host buffers (f16 in, fp32 probabilities out) | " + " | ".join(f"**{k(d[n]['e2e']['value'])}**" for n in d) + " |")
rows.append("| … fp32 in, logits + probabilities out (`full_f32_value`, round 1's budget) | " + " | ".join(k(d[n]['e2e']['full_f32_value']) for n in d) + " |")
rows.append("| CPU port, same step (`--impl reference`), audio-s/s | " + " | ".join(f"{ref[n]['value']:.0f} ({ref[n]['cpu_baseline']['cores']} cores)" for n in d) + " |")
rows.append("| train, 64 windows per GPU (weak), samples/s | " + " | ".join(f"**{k(d[n]['train']['value'])}** ({d[n]['train']['ms_per_step']:.2f} ms)" for n in d) + " |")
rows.append("| … exposed all-reduce, ms | " + " | ".join(f"{d[n]['train']['breakdown_ms']['allreduce_exposed']:.3f}" for n in d) + " |")
rows.append("| train, global batch 64 (strong, the reference's own), samples/s | " + " | ".join(f"{k(d[n]['train_b64']['value'])} ({d[n]['train_b64']['ms_per_step']:.2f} ms)" for n in d) + " |")
rows.append("| 600 s clip → events, clip audio-s/s (ms per clip) | " + " | ".join(f"{k(d[n]['clip']['value'])} ({d[n]['clip']['seconds_per_clip'] * 1e3:.1f})" for n in d) + " |")
rows.append("| validation, 512 windows from host arrays / resident on the device, windows/s | " + " | ".join(f"{k(d[n]['eval']['value'])} / {k(d[n]['eval']['device_resident_value']) if d[n]['eval'].get('device_resident_value') else 'n/a'}" for n in d) + " |")
r1 = d[1]["roofline"]; t1 = d[1]["train"]
rows.append("")
rows.append(f"One GPU: dominant forward family `{r1['kernel']}` {r1['achieved']:.0f} TFLOP/s = **{r1['frac']:.3f}** of the sustained tensor peak "
            f"({r1['kernel_ms_per_step']:.3f} ms of the serial step, {r1['launches_per_step']} launches; HBM figure {r1['hbm']['frac']:.2f}; ncu DRAM traffic {r1['traffic'] / 1e6:.1f} MB per launch cold), "
            f"whole forward **{r1['whole_step_frac']:.3f}** ({r1['whole_step_tflops']:.0f} TFLOP/s); training step {t1['tflops']:.0f} TFLOP/s = {t1['frac_of_tensor_peak']:.3f}, "
            f"dominant family `{t1['roofline']['kernel']}` {t1['roofline']['frac']:.3f}; host-fed training {k(t1['e2e']['value'])} samples/s; CPU port training "
            f"{t1['cpu_baseline']['value']:.1f} samples/s; CPU port batch-1 forward (configs[0]) {d[1]['cpu_baseline']['config1_batch1']['value']:.0f} audio-s/s.  "
            "Round 1 for comparison: forward 204 k / e2e 198 k (1 GPU), 1.61 M / 0.98 M (8 GPUs); train 7.14 k / 52.8 k; no clip, validation or strong-scaling numbers.")
table = "\n".join(rows)
p = os.path.join(ROOT, "DESIGN.md")
s = open(p).read()
if "R02_TABLE" in s:
    s = s.replace("R02_TABLE", "<!-- r02 table -->\n" + table + "\n<!-- /r02 table -->")
else:
    s = re.sub(r"<!-- r02 table -->.*?<!-- /r02 table -->", "<!-- r02 table -->\n" + table.replace("\\", "\\\\") + "\n<!-- /r02 table -->", s, flags=re.S)
open(p, "w").write(s)
print(table)
