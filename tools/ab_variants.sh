#!/bin/bash
# A/B of kernel build variants on one GPU: device-resident timing only (bench.py --no-e2e).
#   tools/ab_variants.sh WORKLOAD lib1.so lib2.so ...      ("in-tree" = p2de_b200/libp2de_b200.so)
# Lines go to gpurun_out/ab_<workload>.jsonl
set -u
W=${1:-S-DMR}; shift
mkdir -p gpurun_out
for lib in "$@"; do
  if [ "$lib" = "in-tree" ]; then unset P2DE_B200_LIB; else export P2DE_B200_LIB="$PWD/$lib"; fi
  timeout 300 python bench.py --workload "$W" --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/ab_err.log | tail -1 | tee -a "gpurun_out/ab_${W}.jsonl"
done
