"""Aggregate an `ncu -i rep --page source --print-source cuda,sass --csv` dump per CUDA source line:
  python profiles/ncu_line_summary.py src.csv [top_n] [kernel-substring]
prints, per kernel section, warp instructions executed and stall samples per source line (top_n lines by instructions)."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(path, newline="")))
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        fpath = r[1]
        fn = rows[i + 1][1] if rows[i + 1][0] == "Function Name" else "?"
        hdr = rows[i + 2]
        i += 3
        c_inst = hdr.index("Instructions Executed")
        c_samp = hdr.index("# Samples")
        c_thr = hdr.index("Thread Instructions Executed")
        per = defaultdict(lambda: [0, 0, 0, ""])
        cur = None
        while i < len(rows) and not (rows[i] and rows[i][0] == "File Path"):
            r = rows[i]
            if len(r) > c_inst:
                if r[0] != "":
                    cur = (fpath.split("/")[-1], int(r[0]))
                    per[cur][3] = r[1].strip()
                elif cur is not None:
                    try:
                        per[cur][0] += int(r[c_inst]); per[cur][1] += int(r[c_samp]); per[cur][2] += int(r[c_thr])
                    except ValueError:
                        pass
            i += 1
        if want and want not in fn:
            continue
        tot_i = sum(v[0] for v in per.values()) or 1
        tot_s = sum(v[1] for v in per.values()) or 1
        print(f"== {fn[:90]}  [{fpath.split('/')[-1]}]  warp instr {tot_i}  samples {tot_s}")
        for (f, ln), v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"  {f}:{ln:<5d} instr {v[0]:>10d} ({100.0 * v[0] / tot_i:5.1f}%)  samples {100.0 * v[1] / tot_s:5.1f}%  lanes {v[2] / max(v[0], 1):4.1f}  {v[3][:100]}")
    else:
        i += 1
