"""Top SASS instructions by stall samples from an ncu report (source page).  python tools/sass_hot.py rep [kernel-index] [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = []; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
rows = list(csv.DictReader(io.StringIO("\n".join(blocks[idx]))))
reasons = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(float(r["# Samples"]) for r in rows)
order = sorted(range(len(rows)), key=lambda i: -float(rows[i]["# Samples"]))[:N]
for i in sorted(order):
    r = rows[i]
    s = float(r["# Samples"])
    top = sorted(((float(r[k]), k[6:]) for k in reasons), reverse=True)[:2]
    print(f"{i:5d} {100*s/tot:5.2f}%  {r['Source'].strip()[:70]:70s} exec={float(r['Instructions Executed'])/1048576:6.2f}  " + ", ".join(f"{k}:{int(v)}" for v, k in top if v > 0))
