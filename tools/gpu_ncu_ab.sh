#!/bin/bash
# tools/gpu_ncu_ab.sh WORKLOAD KERNEL_REGEX name1=lib1.so name2=lib2.so ...: one ncu --set full capture (with SASS source
# counters) of one launch of the matching kernel per library ("in-tree" = the default build) -> gpurun_out/prof_<name>.ncu-rep
W=$1; K=$2; shift 2
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%=*}; lib=${spec#*=}
  if [ "$lib" = "in-tree" ]; then unset P2DE_B200_LIB; else export P2DE_B200_LIB="$PWD/$lib"; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c 1 -o gpurun_out/prof_$name -f \
      python bench.py --workload "$W" --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$name.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
