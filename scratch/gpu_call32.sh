#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/c47_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c47_smoke.log
if ! grep -q "^smoke:" $O/c47_smoke.log; then echo "SMOKE FAILED - stopping"; tail -30 $O/c47_smoke.log; exit 1; fi
for g in 4 6 8 4 8; do
MOPA_SCN_BN_GRID=$g $T 200 python bench.py --gpus 1 --steps 40 --warmup 10 --no-roofline --no-cpu-baseline --no-fp32 > $O/c47_bench.json 2>>$O/c47_err.txt; echo -n "grid=$g "; python scratch/print_bench.py $O/c47_bench.json | cut -c1-200
done
for g in 4 8; do
MOPA_SCN_BN_GRID=$g $T 200 python tools/layer_table.py --out $O/c47_layers_$g.json > $O/c47_layers_$g.log 2>&1; echo "== grid $g"; tail -6 $O/c47_layers_$g.log | grep -E "bn_"
done
$T 900 python -m pytest tests -x -q -m gpu > $O/c47_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c47_tests.log
