#!/bin/bash
# round-2 call 1: mask probe, parity tests with the compacted masked-MMA kernel, bench + layer table, v1 A/B
O=gpurun_out; mkdir -p $O
scratch/bin/umma_mask_probe > $O/c1_mask_probe.txt 2>&1; cat $O/c1_mask_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/c1_tests.log; tail -5 $O/c1_tests.log
timeout 300 python bench.py --no-cpu-baseline > $O/c1_bench.json 2> $O/c1_bench.err; cut -c1-400 $O/c1_bench.json
timeout 300 python tools/layer_table.py --out $O/c1_layers.json > $O/c1_layers.log 2>&1; tail -8 $O/c1_layers.log
MOPA_TC_V1=1 timeout 300 python bench.py --no-cpu-baseline --no-roofline > $O/c1_bench_v1.json 2>/dev/null; cut -c1-200 $O/c1_bench_v1.json
for cfg in "MOPA_TC_CTAS=1" "MOPA_TC_TPC=1" "MOPA_TC_SB=2"; do
  env $cfg timeout 300 python bench.py --no-cpu-baseline --no-roofline --steps 20 > $O/c1_bench_$cfg.json 2>/dev/null; echo $cfg; cut -c1-200 $O/c1_bench_$cfg.json
done
