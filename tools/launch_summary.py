"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time, launches and share per kernel.

    python tools/launch_summary.py gpurun_out/XXX_launches.csv "<header comment>" > profiles/XXX_launches_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    name = re.sub(r"\(.*", "", r[ki])[:90]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
for c in sys.argv[2:]:
    print("#", c)
print(f"# total {total:.1f} us over {sum(cnt.values())} launches\n#   time_us launches share avg_us kernel")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{tot[k]:10.1f} {cnt[k]:5d} {100 * tot[k] / total:5.1f}%  {tot[k] / cnt[k]:8.1f}  {k}")
