#!/bin/bash
# One GPU call at the end of a session: the whole -m gpu suite, smoke(), and a same-box A/B of library builds.
#   tools/final_check.sh [lib ...]       ("in-tree" = p2de_b200/libp2de_b200.so)
mkdir -p gpurun_out
rm -f gpurun_out/final_*.log gpurun_out/final_ab.jsonl
( time timeout 240 python -m pytest tests -m gpu -q -x --ignore=tests/test_gpu_1d_nodewise.py ) > gpurun_out/final_tests.log 2>&1
echo "suite rc=$? $(tail -4 gpurun_out/final_tests.log | head -1)"
( timeout 120 python -m pytest tests/test_gpu_1d_nodewise.py -m gpu -q ) > gpurun_out/final_tests_1d_nodewise.log 2>&1
echo "1d nodewise rc=$? $(tail -1 gpurun_out/final_tests_1d_nodewise.log)"
timeout 90 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/final_smoke.log | cut -c1-160)"
for lib in "$@"; do
  if [ "$lib" = "in-tree" ]; then unset P2DE_B200_LIB; else export P2DE_B200_LIB="$PWD/$lib"; fi
  timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2>>gpurun_out/final_ab_err.log | tail -1 >> gpurun_out/final_ab.jsonl
  echo "$lib $(tail -1 gpurun_out/final_ab.jsonl | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["stage_kernel_ms"], d["value"])')"
done
