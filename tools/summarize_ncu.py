"""Turns gpurun_out ncu artefacts into the small text summaries committed under profiles/.
usage: python tools/summarize_ncu.py <tag> [launches.csv] [full.ncu-rep]"""
import collections
import csv
import subprocess
import sys

tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches.csv"
rep = sys.argv[3] if len(sys.argv) > 3 else None

lines = [l for l in open(launches) if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1000, "us": v, "ms": v * 1000}[row["Metric Unit"]]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write(f"# source: {launches}; total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{t:10.1f} us  {n:4d} launches  {t / tot * 100:5.1f}%  {k}\n")
print(open(f"profiles/{tag}_launches_summary.txt").read())

if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "sm__cycles_active.avg", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit"]
    idx = [(i, h) for i, h in enumerate(hdr) if any(h == w or h.startswith(w + ".") or h.startswith(w + "_") for w in want)]
    with open(f"profiles/{tag}_full_metrics.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source: {rep}\n")
        for r in rows[2:]:
            f.write("----\n")
            for i, h in idx:
                f.write(f"{h} [{units[i]}] = {r[i]}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    kern, cur = [], None
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kern.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    with open(f"profiles/{tag}_full_stalls.txt", "w") as f:
        f.write("# top warp-stall sampling locations (SASS) per captured launch\n")
        for k in kern:
            h = k["hdr"]
            si, so = h.index("# Samples"), h.index("Source")
            sc = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
            total = sum(int(r[si]) for r in k["rows"])
            f.write(f"== {k['name']}  total samples {total}\n")
            for r in sorted(k["rows"], key=lambda r: -int(r[si]))[:12]:
                st = dict(sorted(((h[i], int(r[i])) for i in sc if int(r[i]) > 0), key=lambda kv: -kv[1])[:3])
                f.write(f"  {int(r[si]):6d}  {r[so].strip()[:64]:64s} {st}\n")
    print(open(f"profiles/{tag}_full_stalls.txt").read()[:1500])
