"""Dynamic opcode mix of a kernel from an ncu report's source page (SASS view).

    python tools/sass_mix.py gpurun_out/prof.ncu-rep [kernel-index] [--per N]

Prints warp-level instructions executed per opcode class, stall samples per opcode, divided by N
(e.g. the number of line threads / 32) if given.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def pages(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = {"name": line.split(",", 1)[1].strip('",'), "lines": []}
            blocks.append(cur)
        elif cur is not None:
            cur["lines"].append(line)
    return blocks


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 0
    per = 1.0
    if "--per" in sys.argv:
        per = float(sys.argv[sys.argv.index("--per") + 1])
    b = pages(rep)[idx]
    rd = csv.DictReader(io.StringIO("\n".join(b["lines"])))
    mix, stalls, thr = defaultdict(float), defaultdict(float), defaultdict(float)
    tot = tots = 0.0
    for r in rd:
        src = r["Source"].strip()
        src = re.sub(r"^@!?U?P\w+\s+", "", src)
        op = src.split()[0].rstrip(";")
        base = op.split(".")[0]
        if base in ("LDS", "STS", "LDG", "STG", "IMAD"):
            base = ".".join(op.split(".")[:2]) if base != "IMAD" else ("IMAD.MOV" if "MOV" in op else "IMAD")
        n = float(r["Instructions Executed"]); s = float(r["# Samples"])
        mix[base] += n; stalls[base] += s; thr[base] += float(r["Thread Instructions Executed"])
        tot += n; tots += s
    print(b["name"])
    print(f"warp instructions {tot / per:.1f}  (per unit = /{per:g}); samples {tots:.0f}")
    for k, v in sorted(mix.items(), key=lambda kv: -kv[1]):
        if v / tot < 0.002 and stalls[k] / tots < 0.002:
            continue
        print(f"{k:14s} {v / per:10.1f} {100 * v / tot:6.2f}%   stall samples {100 * stalls[k] / tots:6.2f}%")


if __name__ == "__main__":
    main()
