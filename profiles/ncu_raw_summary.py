"""Key metrics per captured launch from `ncu -i <rep> --page raw --csv`.  usage: python profiles/ncu_raw_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum",
        "l1tex__t_bytes.sum", "smsp__cycles_active.avg"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[ki][:90])
    for w in want:
        if w in hdr:
            print(f"   {w}: {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
