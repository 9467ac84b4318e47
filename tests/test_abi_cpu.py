"""The C-ABI library loads, exports every symbol include/p2de_b200.h declares, validates its
arguments, and refuses to run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import problems as P
from p2de_b200 import lib as plib
from p2de_b200.abi import PackedProblem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "p2de_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p2de_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = plib.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/p2de_b200.h but not exported"
    assert sorted(plib.EXPORTS) == names


def test_config_struct_layout_matches_header():
    """ctypes mirror vs the C struct: same field order, doubles start 8-byte aligned."""
    from p2de_b200.abi import Config
    src = open(os.path.join(ROOT, "include", "p2de_b200.h")).read()
    body = src[src.index("typedef struct p2de_config {"):src.index("} p2de_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"(?:int32_t|int64_t|double)\s+([^;]+);", body):
        fields += [f.strip() for f in decl.split(",")]
    assert fields == [f[0] for f in Config._fields_]
    assert Config.hennemann_a.offset % 8 == 0 and Config.K.offset == 16


def _create(packed):
    L = plib.load()
    h = C.c_void_p()
    rc = L.p2de_create(C.byref(packed.cfg), C.byref(packed.ops), C.byref(packed.geom), C.byref(packed.bc), C.byref(h))
    msg = L.p2de_last_error(None)
    if rc == 0:
        L.p2de_destroy(h)
    return rc, (msg or b"").decode()


def test_argument_validation_and_no_cpu_fallback():
    import torch
    param, rd, md, dd, bc, U0 = P.setup(P.vortex(N=3, K=(3, 3)))
    packed = PackedProblem(param, dd, bc)
    packed.cfg.abi_version = 99
    rc, msg = _create(packed)
    assert rc == -1 and "abi_version" in msg
    packed = PackedProblem(param, dd, bc)
    packed.cfg.Nq = 15
    rc, msg = _create(packed)
    assert rc == -1 and "sizes" in msg
    packed = PackedProblem(param, dd, bc)
    packed.cfg.N = 7
    rc, msg = _create(packed)
    assert rc == -2
    packed = PackedProblem(param, dd, bc)
    rc, msg = _create(packed)
    if torch.cuda.is_available():
        assert rc == 0, msg
    else:
        assert rc == -3 and "no CPU fallback" in msg        # P2DE_ERR_CUDA, loudly


def test_null_handle_calls_do_not_crash():
    L = plib.load()
    assert L.p2de_destroy(None) == 0
    assert L.p2de_synchronize(None) == -1
    assert L.p2de_kernel_launch_count(None) == 0
    out = C.c_double()
    assert L.p2de_rhs(None, 0.0, 0.1, 1, C.byref(out)) == -1


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under p2de_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "p2de_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle." not in txt and "import oracle" not in txt and "p2de_oracle" not in txt, f


def test_ncu_capture_is_of_the_kernels_this_library_holds():
    """profiles/r2_traffic.json (DRAM traffic, FP64 instruction counts of the three stage kernels) is quoted by bench.py only while
    the SASS of exactly those kernels in the built library is what the captured sources compile to; this pins that it is."""
    import json
    import shutil
    import bench
    if not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        pytest.skip("cuobjdump not available")
    with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
        want = json.load(f)["kernel_sass"]
    assert len(want) == 3
    plib.load()
    assert bench.kernel_sass_hashes(plib.SO_PATH, sorted(want)) == want


def test_header_is_plain_c_and_every_entry_point_links(tmp_path):
    """include/p2de_b200.h compiles as C99 (no C++, no torch types in the signatures) and a C program that takes the address of
    every declared entry point links against the built library -- the binding a cgo / ccall / ctypes user gets."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    plib.load()
    names = header_functions()
    src = tmp_path / "abi_link.c"
    src.write_text('#include "p2de_b200.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n  fn_t f[] = {\n'
                   + "".join(f"    (fn_t)&{n},\n" for n in names)
                   + '  };\n  printf("%d %s\\n", (int)(sizeof f / sizeof f[0]), p2de_last_error(0) ? "ok" : "null");\n'
                   '  return p2de_destroy(0);\n}\n')
    exe = tmp_path / "abi_link"
    libdir = os.path.dirname(plib.SO_PATH)
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-l:" + os.path.basename(plib.SO_PATH), "-Wl,-rpath," + libdir]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.split()[0] == str(len(names)), (run.stdout, run.stderr)
