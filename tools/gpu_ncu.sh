# One ncu --set full capture of the four hot-path kernels of one SSP-RK3 step (stage kernel x3, stage-1 combine)
# at the full S-DMR size, plus the launch list of a short bench run.  Outputs under gpurun_out/; tools/summarize_ncu.py
# turns them into the tracked summaries under profiles/.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'stage_kernel_fast|axpy_update_kernel' -s 4 -c 4 -o gpurun_out/prof_stage -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_stage.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out
