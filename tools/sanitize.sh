#!/bin/bash
# compute-sanitizer memcheck / racecheck on small-mesh runs of every kernel family (FAST general / INTERIOR / DEFER,
# generic, Gauss, bounds, 1D).  Summaries go to gpurun_out/sanitize_*.log
mkdir -p gpurun_out
SEL='test_gpu_interior.py::test_interior_ssp33_steps_default_schedule[dmr-N3-64x8] or test_interior_rhs_per_stage[kh-N4-60x6]'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 \
    python -m pytest -q -x -m gpu tests/test_gpu_interior.py -k "dmr-N3-64x8 or kh-N4-60x6 or vortex-N2" \
    > gpurun_out/sanitize_${tool}_fast.log 2>&1; echo "fast $tool rc=$?" >> gpurun_out/sanitize_summary.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 \
    python -m pytest -q -x -m gpu tests/test_gpu_gauss.py tests/test_gpu_bounds.py tests/test_gpu_parity.py -k "N3 or N-3 or 1d or sod or rhs_types or bounds or gauss" --maxfail 1 \
    > gpurun_out/sanitize_${tool}_generic.log 2>&1; echo "generic $tool rc=$?" >> gpurun_out/sanitize_summary.log
done
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitize_*_*.log >> gpurun_out/sanitize_summary.log
cat gpurun_out/sanitize_summary.log
