# Round-2 captures on one GPU (outputs under gpurun_out/, tools/summarize_ncu.py turns them into profiles/r2_*):
#   prof_step.ncu-rep   ncu --set full of the three stage kernels of one SSP-RK3 step at the full S-DMR size (no source import: small)
#   prof_s3.ncu-rep     the stage-3 kernel once more with --import-source on (per-instruction counters for the FP64 counts)
#   launches.csv        launch list of a short bench run
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:'stage_subcell' -s 3 -c 3 -o gpurun_out/prof_step -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stage_subcell_s3' -s 1 -c 1 -o gpurun_out/prof_s3 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_s3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
