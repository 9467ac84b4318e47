#!/usr/bin/env python
"""sha256 of the SASS of selected kernels of a built library (cuobjdump -sass), to show that a source change outside a
kernel left its machine code untouched (profiles/r2_traffic.json carries an ncu capture of stage_subcell_s1/s2/s3<4,8>;
bench.py quotes it only for the kernel sources it was captured from).

usage: tools/kernel_sass_hash.py LIB [LIB2] [--match REGEX]
With two libraries, prints whether every selected function has identical SASS in both.
"""
import hashlib
import re
import subprocess
import sys

DEFAULT = r"stage_subcell_s[123]ILi4ELi8E"


def functions(lib, pattern):
    out = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    res, name, body = {}, None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name is not None:
                res[name] = "\n".join(body)
            name, body = (m.group(1) if re.search(pattern, m.group(1)) else None), []
        elif line.strip().startswith(".......") and name is not None:
            res[name] = "\n".join(body)
            name = None
        elif name is not None:
            body.append(line.rstrip())
    if name is not None:
        res[name] = "\n".join(body)
    return res


def main():
    args = [a for a in sys.argv[1:]]
    pattern = DEFAULT
    if "--match" in args:
        i = args.index("--match")
        pattern = args[i + 1]
        del args[i:i + 2]
    tabs = [functions(lib, pattern) for lib in args]
    for lib, t in zip(args, tabs):
        print(lib)
        for n in sorted(t):
            print(f"  {hashlib.sha256(t[n].encode()).hexdigest()[:16]}  {len(t[n].splitlines()):6d} lines  {n}")
    if len(tabs) == 2:
        same = tabs[0].keys() == tabs[1].keys() and all(tabs[0][n] == tabs[1][n] for n in tabs[0]) and len(tabs[0]) > 0
        print("IDENTICAL" if same else "DIFFERENT")
        return 0 if same else 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
