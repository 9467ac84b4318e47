#!/bin/bash
# One GPU call shaped like the driver's round-end checks: the whole -m gpu suite in ONE pytest process, smoke(), the default
# `python bench.py` (timed), then the S-VORTEX workload (device-resident only).  Results under gpurun_out/end_*.
mkdir -p gpurun_out
rm -f gpurun_out/end_*
( time timeout 150 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/end_tests.log 2>&1
echo "suite rc=$? $(grep -E 'passed|failed|error' gpurun_out/end_tests.log | tail -1)"
( time timeout 60 python __graft_entry__.py smoke ) > gpurun_out/end_smoke.log 2>&1
echo "smoke rc=$? $(grep 'smoke ok' gpurun_out/end_smoke.log | cut -c1-160)"
( time timeout 220 python bench.py ) > gpurun_out/end_bench.json 2> gpurun_out/end_bench_err.log
echo "bench rc=$? $(tail -3 gpurun_out/end_bench_err.log | tr '\n' ' ')"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/end_bench.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("bench", d["value"], d["ms_per_step"], "frac", r["frac"], "traffic", r["traffic"], "|", r["traffic_note"][:90])
    print("e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline", {}).get("value"), "worst", d.get("extra", {}).get("worst_case_value"))
except Exception as e:
    print("bench line unreadable:", repr(e))
PY
timeout 70 python bench.py --workload S-VORTEX --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2>>gpurun_out/end_vortex_err.log | tail -1 > gpurun_out/end_vortex.json
echo "vortex rc=$? $(cut -c1-200 gpurun_out/end_vortex.json)"
