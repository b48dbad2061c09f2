"""profiles/r01_kernel_traffic.json from `ncu -i <rep> --page raw --csv` dumps: per kernel, for the captured launch with the most DRAM traffic: duration, DRAM bytes
(read + write), issue-slot utilisation, tensor-pipe activity, active lanes per instruction and registers.  Later CSVs override earlier ones
per kernel.  usage: python profiles/make_kernel_traffic.py out.json base.json|- raw1.csv [raw2.csv ...]"""
import csv
import json
import re
import sys

out, base, raws = sys.argv[1], sys.argv[2], sys.argv[3:]
doc = json.load(open(base)) if base != "-" else {"source": "", "kernels": {}}
unit_scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


for raw in raws:
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    val = lambda r, k: float(r[col[k]].replace(",", "")) * unit_scale.get(units[col[k]], 1)
    seen = {}
    for r in rows[2:]:
        k = short(r[col["Kernel Name"]])
        e = seen.setdefault(k, {"launches_captured": 0, "time_us_min": 1e30})
        e["launches_captured"] += 1
        t = val(r, "gpu__time_duration.sum")
        e["time_us_min"] = round(min(e["time_us_min"], t), 2)
        dram = int(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
        if dram < e.get("dram_bytes_per_launch_last", -1):
            continue            # keep the figures of the launch with the most DRAM traffic (the full 8-sensor frame, not the host path's per-chunk launches)
        e["time_us_last"] = round(t, 2)
        e["dram_bytes_per_launch_last"] = dram
        e["issue_active_pct"] = round(float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]), 1)
        e["tensor_pipe_pct"] = float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
        e["active_lanes_per_inst"] = round(float(r[col["smsp__thread_inst_executed_per_inst_executed.ratio"]]), 1)
        e["registers"] = r[col["launch__registers_per_thread"]]
        e["capture"] = raw.split("/")[-1]
    doc["kernels"].update(seen)
    doc["source"] = (doc.get("source", "") + "; " if doc.get("source") else "") + f"{raw.split('/')[-1]} (ncu --set full --clock-control none, scripts/prof_run.py, bench inputs)"
json.dump(doc, open(out, "w"), indent=1)
print(out, len(doc["kernels"]), "kernels")
