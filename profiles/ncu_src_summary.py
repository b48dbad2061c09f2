"""Summarise `ncu -i <rep> --page source --csv` output (SASS level, one section per captured launch):
stall-reason shares and the hottest instructions of each section.

usage: ncu -i prof.ncu-rep --page source --csv [--kernel-name regex:<k>] > src.csv
       python profiles/ncu_src_summary.py src.csv [topN] [section_index]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) >= len(cur["hdr"]) - 1:
        cur["data"].append(r)
for si, s in enumerate(sections):
    if only is not None and si != only:
        continue
    hdr, data = s["hdr"], s["data"]
    col = {h: i for i, h in enumerate(hdr)}
    num = lambda r, k: int(float(r[col[k]])) if r[col[k]] not in ("", "-") else 0
    tot = sum(num(r, "# Samples") for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {st: sum(num(r, st) for r in data) for st in stalls}
    print(f"== section {si}: {s['name'][:80]}")
    print("   samples", tot, " warp instructions executed", sum(num(r, "Instructions Executed") for r in data), " SASS lines", len(data))
    print("   stalls:", ", ".join(f"{k[6:]}={v / max(tot, 1):.2f}" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v / max(tot, 1) >= 0.01))
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top]:
        n = num(r, "# Samples")
        ts = sorted(((num(r, st), st[6:]) for st in stalls), reverse=True)[:2]
        print(f"   {n / max(tot, 1):6.3f} exec={r[col['Instructions Executed']]:>9} {r[col['Address']][-5:]} {r[col['Source']].strip()[:70]:70s} {ts[0][1]}:{ts[0][0]} {ts[1][1]}:{ts[1][0]}")
