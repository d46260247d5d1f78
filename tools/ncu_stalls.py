"""Development aid: warp-stall samples of t1_spread_fftx_kernel per phase of the kernel.

Joins `ncu --page source --csv` (SASS rows with sample counts) with `nvdisasm -g` line info of the
in-tree object file and groups the source lines of type1_fused.cuh by the kernel's phases.

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep > profiles/rNN_t1_pass1_stalls.txt
"""
import csv, re, subprocess, sys, tempfile, os
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep = sys.argv[1]
KERNEL = "_ZN2fv21t1_spread_fftx_kernelIfLi9ELi1"
src = (ROOT / "fftvis_b200/csrc/type1_fused.cuh").read_text().split("\n")
def line_of(marker):
    return next(i + 1 for i, l in enumerate(src) if marker in l)
marks = [("fft butterflies + stages", 1), ("kernel prologue + strip clear", line_of("t1_spread_fftx_kernel(T1SpreadArgs<T> a) {")),
         ("source scan", line_of("which sources' w-row footprints touch this strip")),
         ("records + kernel samples (fill)", line_of("flush: evaluate kernels densely")),
         ("spread (column segments)", line_of("if (use_seg) {")),
         ("spread (row blocks, small grids)", line_of("} else if (warp * ((rows + nwarps - 1) / nwarps) < rows) {")),
         ("row FFT call + write-out", line_of("rows of all products are `pitch` apart"))]
def phase(loc):
    if loc is None:
        return "unmapped"
    f, l = loc
    if f == "common.cuh":
        return "common.cuh (cmul in the FFT twiddles, es_kernel in fill)"
    if f != "type1_fused.cuh":
        return f
    name = marks[0][0]
    for n, start in marks:
        if l >= start:
            name = n
    return name

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "fftvis_b200/csrc/_obj/type1_fused.o")], cwd=td, check=True, capture_output=True)
    cubin = next(Path(td).glob("*.cubin"))
    sass = subprocess.run(["nvdisasm", "-g", str(cubin)], capture_output=True, text=True, check=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text." + KERNEL))
end = next(i for i, l in enumerate(sass) if l.startswith(".text.") and i > start)
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
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:t1_spread"],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
base, tot, totex, ph = None, 0, 0, {}
for r in rows[hi + 1:]:
    try:
        a = int(r[0], 16)
    except (ValueError, IndexError):
        continue
    base = a if base is None else base
    s, e = int(r[isamp] or 0), int(r[iex] or 0)
    d = ph.setdefault(phase(amap.get(a - base)), {"samples": 0, "inst": 0})
    d["samples"] += s; d["inst"] += e; tot += s; totex += e
    for c in stalls:
        d[c] = d.get(c, 0) + int(r[hdr.index(c)] or 0)
print(f"t1_spread_fftx_kernel<float, 9>: {tot} warp samples, {totex} warp instructions ({rep})")
print(f"{'phase':58s} {'samples':>8s} {'%':>6s} {'inst %':>7s}  top stall reasons")
for p, d in sorted(ph.items(), key=lambda kv: -kv[1]["samples"]):
    top = sorted(((c[6:], d[c]) for c in stalls), key=lambda kv: -kv[1])[:4]
    print(f"{p:58s} {d['samples']:8d} {100 * d['samples'] / tot:6.1f} {100 * d['inst'] / max(totex, 1):7.1f}  "
          + ", ".join(f"{k} {v}" for k, v in top))
