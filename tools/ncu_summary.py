"""Development aid: transpose `ncu -i <rep> --page raw --csv` into metric rows (one column per launch),
and optionally aggregate the warp-stall samples of one kernel per source line.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_summary.csv
"""
import csv, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
    for j, name in enumerate(hdr):
        if name in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC",
                    "Section Name", "Metric Name", "Metric Unit", "Metric Value"):
            continue
        w.writerow([name, units[j]] + [r[j] for r in data])
print("wrote", out, len(hdr), "metrics x", len(data), "launches")
