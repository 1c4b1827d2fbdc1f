#!/bin/bash
# round-2 call 2: tile-rulebook conv kernel (v3): parity tests, bench, layer table, config sweep
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/c2_tests.log; tail -5 $O/c2_tests.log
timeout 300 python bench.py --no-cpu-baseline > $O/c2_bench.json 2> $O/c2_bench.err; cut -c1-300 $O/c2_bench.json; tail -3 $O/c2_bench.err
timeout 300 python tools/layer_table.py --out $O/c2_layers.json > $O/c2_layers.log 2>&1; tail -8 $O/c2_layers.log
for cfg in "MOPA_TC_CTAS=1" "MOPA_TC_TPC=1" "MOPA_TC_TPC=2" "MOPA_TC_SB=2" "MOPA_TC_NA=1"; do
  env $cfg timeout 300 python bench.py --no-cpu-baseline --no-roofline --steps 20 > $O/c2_bench_$cfg.json 2>/dev/null; echo $cfg; cut -c1-200 $O/c2_bench_$cfg.json
done
