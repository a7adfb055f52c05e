"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python scripts/summarise_launches.py gpurun_out/launches.csv [steps] [skip_first_n_launches]
"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("espic::", "").replace("void ", "")
    rows.append((name, us))
rows = rows[skip:]
tot = OrderedDict()
for name, us in rows:
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += us
total = sum(v[1] for v in tot.values())
print(f"launches {len(rows)}  total {total:.1f} us" + (f"  = {total / steps:.1f} us/step over {steps} steps" if steps else ""))
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    per = f"  per step {us / steps:9.1f} us" if steps else ""
    print(f"{name[:60]:60s} launches {n:5d}{per}  mean {us / n:9.1f} us  share {100 * us / total:5.1f} %")
