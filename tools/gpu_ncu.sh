# Round-2 captures on one GPU (outputs under gpurun_out/, tools/summarize_ncu.py turns them into profiles/r2_*).
# gpurun brings back at most 64 MiB per call, so the two reports are made by separate calls:
#   tools/gpu_ncu.sh step   prof_step.ncu-rep: ncu --set full of the three stage kernels of one SSP-RK3 step at the full S-DMR size
#   tools/gpu_ncu.sh s3     prof_s3.ncu-rep: the stage-3 kernel with --import-source on (per-instruction counters), + launches.csv
mkdir -p gpurun_out
if [ "$1" = "step" ]; then
  ncu --set full --clock-control none -k regex:'stage_subcell' -s 3 -c 3 -o gpurun_out/prof_step -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_step.log 2>&1
else
  ncu --set full --clock-control none --import-source on -k regex:'stage_subcell_s3' -s 1 -c 1 -o gpurun_out/prof_s3 -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_s3.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
