#!/usr/bin/env python
"""Turns gpurun_out/prof_step.ncu-rep + prof_s3.ncu-rep + launches.csv (tools/gpu_ncu.sh) into the tracked summaries:

    profiles/r2_ncu_full_raw_S-DMR.csv              ncu --page raw --csv of the three stage kernels of one step
    profiles/r2_ncu_details_excerpt.txt             selected lines of --page details
    profiles/r2_launches_bench_steps2_warmup1.csv   launch list (gpu__time_duration.sum per launch)
    profiles/r2_traffic.json                        DRAM bytes, pipe utilisation, FP64 instruction counts per launch, stamped
                                                    with the hash of the kernel sources (bench.py quotes it only on a match)

Runs on the CPU box (ncu -i needs no GPU)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
NODES = 4096 * 1024 * 16     # S-DMR: DOF-updates per stage


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", os.path.join(OUT, rep), *args], capture_output=True, text=True, check=True).stdout


def main():
    from bench import source_hash
    raw = ncu("prof_step.ncu-rep", "--page", "raw", "--csv")
    with open(os.path.join(PROF, "r2_ncu_full_raw_S-DMR.csv"), "w") as f:
        f.write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}

    def val(r, name):
        return float(r[col[name]].replace(",", ""))

    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}
    out, total = {}, 0.0
    for r in data:
        m = re.search(r"stage_subcell_(s1|s2|s3|rt)", r[col["Kernel Name"]])
        if not m or m.group(1) in out:
            continue
        nm = m.group(1)
        rd = val(r, "dram__bytes_read.sum") * scale[units[col["dram__bytes_read.sum"]]]
        wr = val(r, "dram__bytes_write.sum") * scale[units[col["dram__bytes_write.sum"]]]
        dur_ms = val(r, "gpu__time_duration.sum") * tscale[units[col["gpu__time_duration.sum"]]]

        def opt(name, f=1.0):
            return f * val(r, name) if name in col else None
        out["stage_subcell_" + nm] = {
            "dram_read_bytes": rd, "dram_write_bytes": wr, "duration_ms_under_ncu": round(dur_ms, 4),
            "registers_per_thread": int(val(r, "launch__registers_per_thread")),
            "shared_mem_per_block_bytes": opt("launch__shared_mem_per_block_dynamic", scale.get(units[col["launch__shared_mem_per_block_dynamic"]], 1.0)) if "launch__shared_mem_per_block_dynamic" in col else None,
            "fp64_pipe_cycles_active_pct": opt("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_slots_busy_pct": opt("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "lsu_data_pipe_wavefronts_pct": opt("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "shared_wavefronts": opt("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            "dram_throughput_pct": opt("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_per_sm": opt("sm__warps_active.avg.per_cycle_active"),
            "local_loads": opt("sass__inst_executed_local_loads"), "local_stores": opt("sass__inst_executed_local_stores")}
        total += rd + wr
    # FP64 thread-instruction counts of the stage-3 kernel from the source page
    srows = list(csv.reader(io.StringIO(ncu("prof_s3.ncu-rep", "--page", "source", "--csv", "--print-source", "sass"))))
    hdr_i = [i for i, r in enumerate(srows) if r and r[0] == "Address"][0]
    ci = {n: i for i, n in enumerate(srows[hdr_i])}
    f = {"dfma": 0.0, "dmul": 0.0, "dadd": 0.0, "dsetp": 0.0, "all": 0.0, "lds": 0.0, "sts": 0.0, "mufu": 0.0}
    for r in srows[hdr_i + 1:]:
        if len(r) <= ci["Predicated-On Thread Instructions Executed"] or not r[0].startswith("0x"):
            continue
        n = float(r[ci["Predicated-On Thread Instructions Executed"]])
        f["all"] += n
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[ci["Source"]])
        op = m.group(2).lower() if m else ""
        if op in f:
            f[op] += n
    f = {k: v / NODES for k, v in f.items()}
    out["per_step_total_bytes"] = total
    out["per_stage_total_bytes"] = total / 3
    out["fp64"] = {"_comment": "stage_subcell_s3: predicated-on thread instructions per DOF-update (67.1 M nodes) from the ncu source page",
                   "dfma_per_dof_update": round(f["dfma"], 2), "dmul_per_dof_update": round(f["dmul"], 2), "dadd_per_dof_update": round(f["dadd"], 2),
                   "flops_per_dof_update": round(2 * f["dfma"] + f["dmul"] + f["dadd"], 2),
                   "lane_ops_per_dof_update": round(f["dfma"] + f["dmul"] + f["dadd"], 2),
                   "dsetp_per_dof_update": round(f["dsetp"], 2), "mufu_per_dof_update": round(f["mufu"], 2),
                   "lds_per_dof_update": round(f["lds"], 2), "sts_per_dof_update": round(f["sts"], 2),
                   "thread_instructions_per_dof_update": round(f["all"], 1)}
    js = {"_comment": "Per-launch numbers from ONE ncu --set full --clock-control none capture (profiles/r2_ncu_full_raw_S-DMR.csv, made by "
                      "tools/gpu_ncu.sh + tools/summarize_ncu.py), workload S-DMR (4096x1024, N=3, subcell), 1 B200: the three stage kernels "
                      "of one SSP-RK3 step.  bench.py quotes per_stage_total_bytes and the fp64 block only while source_hash matches the "
                      "kernel sources it runs.",
          "source_hash": source_hash(), "S-DMR": out}
    with open(os.path.join(PROF, "r2_traffic.json"), "w") as fjs:
        json.dump(js, fjs, indent=1)
    det = ncu("prof_step.ncu-rep", "--page", "details")
    keep = ("stage_subcell", "Memory Throughput", "DRAM Throughput", "Duration", "L2 Cache Throughput", "Compute (SM) Throughput", "L1/TEX",
            "Executed Ipc Active", "Issue Slots Busy", "FP64", "fused", "Block Size", "Grid Size", "Registers Per Thread", "Dynamic Shared Memory Per Block",
            "Block Limit", "Theoretical Occupancy", "Achieved Occupancy", "Warp Cycles Per Issued", "No Eligible", "Eligible Warps")
    with open(os.path.join(PROF, "r2_ncu_details_excerpt.txt"), "w") as fd:
        for line in det.splitlines():
            if any(k in line for k in keep):
                fd.write(line.rstrip()[:160] + "\n")
    src = os.path.join(OUT, "launches.csv")
    if os.path.exists(src):
        with open(src) as fi, open(os.path.join(PROF, "r2_launches_bench_steps2_warmup1.csv"), "w") as fo:
            for line in fi:
                if line.startswith('"'):
                    fo.write(line)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
