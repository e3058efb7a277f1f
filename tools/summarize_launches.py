"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the LAST step of bench.py.
    python tools/summarize_launches.py gpurun_out/launches.csv <launches per step> > profiles/rNN_launch_list_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    started = False
    for r in csv.reader(f):
        if not started:
            started = r[:1] == ["ID"]
            hdr = r
            continue
        if len(r) == len(hdr):
            rows.append(dict(zip(hdr, r)))
rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
def ms(r):
    v = float(r["Metric Value"].replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r["Metric Unit"]]
names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("vl::", "").replace("void ", "") for r in rows]
# last step = everything after the last adamw launch but one
adam = [i for i, n in enumerate(names) if "adamw" in n]
lo, hi = (adam[-2] + 1, adam[-1] + 1) if len(adam) >= 2 else (0, len(rows))
agg = OrderedDict()
for i in range(lo, hi):
    a = agg.setdefault(names[i], [0.0, 0])
    a[0] += ms(rows[i])
    a[1] += 1
tot = sum(a[0] for a in agg.values())
print(f"last step: {hi - lo} kernel launches, {tot:.2f} ms summed device time (serialised, cold-cache: compare SHARES)")
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {t:8.3f} ms  {100 * t / tot:4.1f}%  n={c:5d}  {n[:90]}")
