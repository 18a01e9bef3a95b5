"""dram__bytes_read.sum + dram__bytes_write.sum per launch of the first kernel in an ncu report, as one JSON object line:
python tools/ncu_traffic.py <report.ncu-rep> <family name>"""
import csv, json, subprocess, sys
rep, name = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = 0.0
for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    i = hdr.index(k)
    tot += float(r[i].replace(",", "")) * mult.get(units[i], 1)
print(json.dumps({name: int(tot)}))
