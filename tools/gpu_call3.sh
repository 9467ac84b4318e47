mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=line 2>&1 | grep -v "^$" | cut -c1-300 | tail -12 > gpurun_out/t_all6.log
tools/ab_variants.sh S-DMR in-tree
P2DE_NO_DEFER=1 tools/ab_variants.sh S-DMR in-tree
tools/ab_variants.sh S-KH in-tree
P2DE_NO_DEFER=1 tools/ab_variants.sh S-KH in-tree
tail -5 gpurun_out/t_all6.log
