#!/usr/bin/env python
"""Turns gpurun_out/prof_stage.ncu-rep + gpurun_out/launches.csv (tools/gpu_ncu.sh) into the tracked summaries:

    profiles/r1_ncu_full_raw_S-DMR.csv          ncu --page raw --csv of the capture
    profiles/r1_ncu_details_excerpt.txt         selected lines of --page details
    profiles/r1_launches_bench_steps2_warmup1.csv   launch list (gpu__time_duration.sum per launch)
    profiles/r1_traffic.json                    DRAM bytes / FP64 instruction counts per launch (read by bench.py)

Runs on the CPU box (ncu -i needs no GPU)."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REP = os.path.join(ROOT, "gpurun_out", "prof_stage.ncu-rep")
PROF = os.path.join(ROOT, "profiles")
NODES = 4096 * 1024 * 16     # S-DMR: DOF-updates per stage


def ncu(*args):
    return subprocess.run(["ncu", "-i", REP, *args], capture_output=True, text=True, check=True).stdout


def main():
    raw = ncu("--page", "raw", "--csv")
    with open(os.path.join(PROF, "r1_ncu_full_raw_S-DMR.csv"), "w") as f:
        f.write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, data = rows[0], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}

    def val(r, name):
        return float(r[col[name]].replace(",", ""))

    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}
    out = {}
    total = 0.0
    for r in data:
        kn = r[col["Kernel Name"]]
        rd = val(r, "dram__bytes_read.sum") * scale[rows[1][col["dram__bytes_read.sum"]]]
        wr = val(r, "dram__bytes_write.sum") * scale[rows[1][col["dram__bytes_write.sum"]]]
        # the launches of one step: stage 1 (reads Uq only), stage 2 (its own instantiation: forms U1 while loading), stage 3
        if "axpy_update_kernel" in kn:
            nm = "axpy_update_kernel_stage1"
        elif "stage_kernel_fast_defer" in kn:
            nm = "stage_kernel_fast_defer_stage2"
        elif "stage_kernel_fast" in kn:
            nm = "stage_kernel_fast_stage1" if rd < 3.0e9 else "stage_kernel_fast_stage3"
        else:
            continue
        if nm in out:
            continue
        dur_ms = val(r, "gpu__time_duration.sum") * tscale[rows[1][col["gpu__time_duration.sum"]]]

        def opt(name, f=1.0):
            return f * val(r, name) if name in col else None
        out[nm] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "duration_ms_under_ncu": round(dur_ms, 4),
                   "registers_per_thread": int(val(r, "launch__registers_per_thread")),
                   "fp64_pipe_cycles_active_pct": opt("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                   "issue_slots_busy_pct": opt("sm__inst_issued.avg.pct_of_peak_sustained_active"),
                   "dram_throughput_pct": opt("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                   "warps_active_per_sm": opt("smsp__warps_active.avg.per_cycle_active", 4.0)}
        total += rd + wr
    # FP64 thread-instruction counts of the stage-2 kernel from the source page (the third captured launch)
    import re
    srows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-kernel-base", "function"))))
    kernels, cur = [], None
    for r in srows:
        if r and r[0] == "Kernel Name":
            cur = {"rows": [], "name": r[1]}; kernels.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    k2 = [k for k in kernels if "stage_kernel_fast" in k["name"] and "defer" not in k["name"]][0]
    for kk in kernels:      # prefer a stage-3 launch (plain instantiation, SSP combine fused): the one with resW traffic
        if "stage_kernel_fast" in kk["name"] and "defer" not in kk["name"]:
            k2 = kk
    ci = {n: i for i, n in enumerate(k2["hdr"])}
    f = {"dfma": 0.0, "dmul": 0.0, "dadd": 0.0, "dsetp": 0.0, "all": 0.0}
    for r in k2["rows"]:
        n = float(r[ci["Predicated-On Thread Instructions Executed"]])
        f["all"] += n
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ci["Source"]])
        op = m.group(2).lower() if m else ""
        if op in f:
            f[op] += n
    f = {k: v / NODES for k, v in f.items()}
    out["per_step_total_bytes"] = total
    out["per_stage_total_bytes"] = total / 3
    out["fp64"] = {"_comment": "stage kernel (last captured launch of the plain instantiation): predicated-on thread instructions per DOF-update (67.1 M nodes) from the ncu source page",
                   "dfma_per_dof_update": round(f["dfma"], 2), "dmul_per_dof_update": round(f["dmul"], 2), "dadd_per_dof_update": round(f["dadd"], 2),
                   "flops_per_dof_update": round(2 * f["dfma"] + f["dmul"] + f["dadd"], 2),
                   "lane_ops_per_dof_update": round(f["dfma"] + f["dmul"] + f["dadd"], 2),
                   "dsetp_per_dof_update": round(f["dsetp"], 2), "thread_instructions_per_dof_update": round(f["all"], 1)}
    js = {"_comment": "Per-launch numbers from ONE ncu --set full --clock-control none capture (profiles/r1_ncu_full_raw_S-DMR.csv, made by "
                      "tools/gpu_ncu.sh + tools/summarize_ncu.py), workload S-DMR (4096x1024, N=3, subcell), 1 B200: the four hot-path launches of "
                      "one SSP-RK3 step (stage kernel; stage kernel forming the stage-1 combine while loading; stage kernel). bench.py reads "
                      "per_stage_total_bytes and the fp64 block.",
          "S-DMR": out}
    with open(os.path.join(PROF, "r1_traffic.json"), "w") as fjs:
        json.dump(js, fjs, indent=1)
    det = ncu("--page", "details")
    keep = ("stage_kernel_fast", "axpy_update_kernel", "Memory Throughput", "DRAM Throughput", "Duration", "L2 Cache Throughput", "Compute (SM) Throughput",
            "Executed Ipc Active", "Issue Slots Busy", "FP64", "fused", "Block Size", "Grid Size", "Registers Per Thread", "Dynamic Shared Memory Per Block",
            "Block Limit", "Theoretical Occupancy", "Achieved Occupancy", "Warp Cycles Per Issued", "No Eligible", "Eligible Warps")
    with open(os.path.join(PROF, "r1_ncu_details_excerpt.txt"), "w") as fd:
        for line in det.splitlines():
            if any(k in line for k in keep):
                fd.write(line.rstrip()[:160] + "\n")
    src = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if os.path.exists(src):
        with open(src) as fi, open(os.path.join(PROF, "r1_launches_bench_steps2_warmup1.csv"), "w") as fo:
            for line in fi:
                if line.startswith('"'):
                    fo.write(line)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
