#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/c53_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c53_smoke.log
if ! grep -q "^smoke:" $O/c53_smoke.log; then echo "SMOKE FAILED - stopping"; tail -30 $O/c53_smoke.log; exit 1; fi
$T 600 python -m pytest tests -x -q -m gpu > $O/c53_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c53_tests.log
$T 200 python tools/layer_table.py --out $O/c53_layers.json > $O/c53_layers.log 2>&1; grep -E "subm +1 +16" $O/c53_layers.log; tail -6 $O/c53_layers.log 
for v in 1 2; do
$T 200 python bench.py --gpus 1 --steps 40 --warmup 10 --no-roofline --no-cpu-baseline --no-fp32 > $O/c53_bench.json 2>>$O/c53_err.txt; python scratch/print_bench.py $O/c53_bench.json | cut -c1-200
done
