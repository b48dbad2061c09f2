"""Summarise an ncu --csv launch list (gpu__time_duration.sum etc.) of scripts/prof_icp.py: python scripts/icp_launch_table.py file.csv"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
d = defaultdict(dict)
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    d[(int(r[0]), r[h.index("Kernel Name")].split("(")[0])][r[h.index("Metric Name")]] = float(r[h.index("Metric Value")].replace(",", ""))
agg = defaultdict(list)
for (i, k), v in sorted(d.items()):
    agg[k].append(v)
for k, vs in agg.items():
    if not any(x in k for x in ("match", "stats", "sums", "reduce")):
        continue
    vs = vs[1:] if len(vs) > 1 else vs          # skip the first iteration (no seeds)
    n = len(vs)
    m = lambda key: sum(v.get(key, 0.0) for v in vs) / n
    print(f"{k:24s} n={n} dur {m('gpu__time_duration.sum') / 1000:7.1f} us  instr {m('smsp__inst_executed.sum') / 1e6:6.2f} M  issue {m('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} %  "
          f"smsp active {100 * m('smsp__cycles_active.avg') / max(m('sm__cycles_elapsed.max'), 1):5.1f} % of elapsed")
