#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/c37_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c37_smoke.log
if ! grep -q "^smoke:" $O/c37_smoke.log; then echo "SMOKE FAILED - stopping"; tail -30 $O/c37_smoke.log; exit 1; fi
for rep in 1 2; do
$T 200 python bench.py --steps 40 --warmup 10 --no-fp32 --no-roofline --no-cpu-baseline > $O/c37_bench_$rep.json 2>>$O/c37_err.txt; python scratch/print_bench.py $O/c37_bench_$rep.json
done
$T 200 python tools/layer_table.py --out $O/c37_layers.json > $O/c37_layers.log 2>&1; grep -E "conv_dweight +subm +1 " $O/c37_layers.log; tail -6 $O/c37_layers.log
$T 900 python -m pytest tests -x -q -m gpu > $O/c37_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/c37_tests.log
