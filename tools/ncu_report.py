"""Print key metrics and top stall lines of an ncu report. usage: python tools/ncu_report.py rep [ntop]"""
import csv, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]
def col(n): return hdr.index(n)
keys = ['gpu__time_duration.sum','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__cycles_active.avg','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__waves_per_multiprocessor','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_pipe_xu.sum','launch__occupancy_per_block_size']
for r in rows[2:3]:
    print(r[col('Kernel Name')][:70], 'grid', r[col('Grid Size')], 'block', r[col('Block Size')])
    for k in keys:
        if k in hdr: print(f"   {k} = {r[col(k)]} {rows[1][col(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(src.splitlines()):
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'hdr': None, 'rows': []}; kern.append(cur)
    elif cur is not None and cur['hdr'] is None: cur['hdr'] = r
    elif cur is not None: cur['rows'].append(r)
k = kern[0]; h = k['hdr']; si = h.index('# Samples'); so = h.index('Source')
sc = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r[si]) for r in k['rows']); agg = {}
for r in k['rows']:
    for i in sc: agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print('total samples', tot, dict(sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
for idx, r in sorted(enumerate(k['rows']), key=lambda ir: -int(ir[1][si]))[:ntop]:
    st = dict(sorted(((h[i], int(r[i])) for i in sc if int(r[i]) > 0), key=lambda kv: -kv[1])[:2])
    print(f"  {int(r[si]):6d} @{idx:5d} {r[so].strip()[:58]:58s} {st}")
