"""Stall samples and executed warp instructions per CUDA source line of one kernel of an ncu report.
python tools/src_phases.py rep [kernel-index] [min-percent]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
# the output is a sequence of (File Path, Function Name, table) groups; keep those of the idx-th distinct function
groups, cur = [], None
for line in out.splitlines():
    if line.startswith('"File Path"'):
        cur = {"file": line.split(",", 1)[1].strip('"'), "lines": []}
        groups.append(cur)
    elif line.startswith('"Function Name"') and cur is not None:
        cur["func"] = line.split(",", 1)[1].strip('"')
    elif cur is not None:
        cur["lines"].append(line)
funcs = []
for g in groups:
    if g.get("func") not in funcs: funcs.append(g.get("func"))
f = funcs[idx]
rows = []
for g in groups:
    if g.get("func") != f: continue
    rd = csv.reader(io.StringIO("\n".join(g["lines"])))
    hdr = next(rd)
    iS, iN = hdr.index("# Samples"), hdr.index("Instructions Executed")
    for rr in rd:
        r = {"Line No": rr[0], "Source": rr[1], "# Samples": rr[iS], "Instructions Executed": rr[iN]}
        if not rr[0]: continue
        try: rows.append((g["file"].split("/")[-1], int(r["Line No"]), r["Source"].strip(), float(r["# Samples"]), float(r["Instructions Executed"])))
        except (ValueError, KeyError): pass
tot = sum(r[3] for r in rows); toti = sum(r[4] for r in rows)
print(f, "samples", tot)
for fn, ln, src, s, n in rows:
    if 100 * s / tot >= minp:
        print(f"{fn}:{ln:4d} {100*s/tot:5.2f}% inst {100*n/toti:5.2f}%  {src[:110]}")
