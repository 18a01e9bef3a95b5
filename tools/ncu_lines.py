"""Stall samples of an ncu report aggregated per CUDA source line (the CSV source page only lists SASS).
The k-th SASS row of the report is the k-th instruction of the kernel in `nvdisasm -g` of the in-tree library, whose
'//## File "...", line N' annotations give the line.   usage: python tools/ncu_lines.py rep.ncu-rep kernel_substring [ntop] [mangled_section_substring]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
section = sys.argv[4] if len(sys.argv) > 4 else kern   # e.g. block_fused_kernelILi128ELb0 to pick one template instance
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("A2M_PROFILE_LIB") or os.path.join(root, "audio-to-midi_b200", "_build", "libaudio2midi_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "modelutil" not in f][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instructions of the kernel, in order, with the innermost source line
lines, cur, inside = [], None, False
for ln in sass:
    if ln.startswith("\t.section\t.text."):
        inside = section in ln
        continue
    if ln.startswith("\t.section"):
        inside = False
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        if "inlined at" not in m.group(3) or cur is None:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
k, cur = [], None
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        k.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
kk = [x for x in k if kern in x["name"]][0]
h = kk["hdr"]
si = h.index("# Samples")
sc = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
print(f"{kk['name'][:60]}: {len(kk['rows'])} SASS rows in the report, {len(lines)} in the disassembly")
agg = {}
for idx, r in enumerate(kk["rows"]):
    key = lines[idx] if idx < len(lines) and lines[idx] else ("?", 0)
    a = agg.setdefault(key, {"n": 0, "st": {}})
    a["n"] += int(r[si])
    for i in sc:
        if int(r[i]):
            a["st"][h[i]] = a["st"].get(h[i], 0) + int(r[i])
tot = sum(a["n"] for a in agg.values())
srcs = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:ntop]:
    if f not in srcs:
        p = os.path.join(root, "audio-to-midi_b200", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][l - 1].strip()[:70] if 0 < l <= len(srcs[f]) else ""
    st = dict(sorted(a["st"].items(), key=lambda kv: -kv[1])[:2])
    print(f"{a['n']:6d} {100 * a['n'] / tot:5.1f}%  {f}:{l:<4d} {text:70s} {st}")
