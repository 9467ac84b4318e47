mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
(time python -m pytest tests/test_gpu_bounds.py -q --maxfail=30 -m gpu 2>&1 | tail -40) > gpurun_out/t_bounds.log 2>&1
(time python -m pytest tests -q --maxfail=20 -m gpu --deselect tests/test_gpu_bounds.py 2>&1 | tail -40) > gpurun_out/t_all.log 2>&1
tools/ab_variants.sh S-DMR build_variants/v0.so in-tree build_variants/pf148.so build_variants/pf592.so build_variants/pf2368.so in-tree
tools/ab_variants.sh S-KH build_variants/v0.so in-tree build_variants/pf592.so
tail -5 gpurun_out/t_bounds.log gpurun_out/t_all.log
