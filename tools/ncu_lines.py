"""Development aid: warp-stall samples and executed instructions per SOURCE LINE of one kernel.

Joins `ncu --page source --csv` (SASS rows with sample counts) with `nvdisasm -g` line info of the
in-tree object file.

    python tools/ncu_lines.py <report.ncu-rep> <object file under fftvis_b200/csrc/_obj> <mangled-name prefix> [launch index]
"""
import csv, re, subprocess, sys, tempfile
from collections import defaultdict
from pathlib import Path

rep, obj, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
launch = sys.argv[4] if len(sys.argv) > 4 else "0"
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=td, check=True, capture_output=True)
    cubin = next(Path(td).glob("*.cubin"))
    sass = subprocess.run(["nvdisasm", "-g", str(cubin)], capture_output=True, text=True, check=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text." + kernel))
end = next((i for i, l in enumerate(sass) if l.startswith(".text.") and i > start), len(sass))
cur, amap = None, {}
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if "inlined" not in m.group(3):
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+\S", l)
    if m:
        amap[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", launch, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Address" in r)
h = rows[hi]
ia, isamp, iex = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
agg = defaultdict(lambda: [0, 0])
base = None
for r in rows[hi + 1:]:
    if len(r) != len(h) or r[ia] == "Address":
        continue
    a = int(r[ia], 16)
    base = a if base is None else base
    loc = amap.get(a - base)
    agg[loc][0] += int(r[isamp] or 0)
    agg[loc][1] += int(r[iex] or 0)
ts, te = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"total samples {ts}, warp instructions {te}")
srcs = {}
for loc, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    text = ""
    if loc:
        f = next(Path("fftvis_b200/csrc").glob(loc[0]), None)
        if f:
            srcs.setdefault(f, f.read_text().split("\n"))
            text = srcs[f][loc[1] - 1].strip()[:90]
    print(f"{100 * s / ts:5.1f}% samples {100 * e / te:5.1f}% instr  {loc}  {text}")
