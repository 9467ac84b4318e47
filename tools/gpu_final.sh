# round-end style verification on one GPU: GPU tests, smoke, ncu capture, bench lines of the three workloads
mkdir -p gpurun_out
(time python -m pytest tests -q -m gpu -x 2>&1 | tail -4) > gpurun_out/final_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
bash tools/gpu_ncu.sh > /dev/null 2>&1
python bench.py --steps 10 --warmup 3 2>gpurun_out/final_bench_err.log | tail -1 > gpurun_out/bench_S-DMR_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/final_bench_err.log | tail -1 > gpurun_out/bench_S-DMR_reference.json
python bench.py --workload S-KH --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/final_bench_err.log | tail -1 > gpurun_out/bench_S-KH_1gpu.json
python bench.py --workload S-KH-gauss --steps 6 --warmup 3 --no-cpu-baseline 2>>gpurun_out/final_bench_err.log | tail -1 > gpurun_out/bench_S-KH-gauss_1gpu.json
cat gpurun_out/final_tests.log gpurun_out/final_smoke.log
