"""Development aid: print the key metrics + top stall reasons of every launch in an ncu report."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg"]
def f(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
for r in rows[2:]:
    print("----")
    for w in want:
        if w in h:
            print(f"{w:85s} {r[h.index(w)]} {rows[1][h.index(w)]}")
    st = [(h[i], r[i]) for i in range(len(h)) if "smsp__average_warps_issue_stalled" in h[i] and h[i].endswith("_per_issue_active.ratio")]
    for k, v in sorted(st, key=lambda kv: -f(kv[1]))[:8]:
        print(f"   {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v}")
