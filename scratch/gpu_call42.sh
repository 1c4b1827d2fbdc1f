#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/c52_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c52_smoke.log
if ! grep -q "^smoke:" $O/c52_smoke.log; then echo "SMOKE FAILED - stopping"; tail -30 $O/c52_smoke.log; exit 1; fi
$T 600 python -m pytest tests -x -q -m gpu > $O/c52_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c52_tests.log
$T 200 python tools/layer_table.py --out $O/c52_layers.json > $O/c52_layers.log 2>&1; grep -E "conv_dweight +subm +1 " $O/c52_layers.log; tail -6 $O/c52_layers.log | grep dweight
for v in 1 2; do
$T 200 python bench.py --gpus 1 --steps 40 --warmup 10 --no-roofline --no-cpu-baseline --no-fp32 > $O/c52_bench.json 2>>$O/c52_err.txt; python scratch/print_bench.py $O/c52_bench.json | cut -c1-200
done
