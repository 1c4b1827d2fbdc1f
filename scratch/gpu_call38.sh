#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/c51_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/c51_smoke.log
if ! grep -q "^smoke:" $O/c51_smoke.log; then echo "SMOKE FAILED - stopping"; tail -30 $O/c51_smoke.log; exit 1; fi
$T 600 python -m pytest tests -x -q -m gpu > $O/c51_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c51_tests.log
for cfg in "X=0" "MOPA_TC_HALF_SMALL=1"; do
tag="${cfg// /_}"
env $cfg $T 200 python tools/layer_table.py --out $O/c51_layers_$tag.json > $O/c51_layers_$tag.log 2>&1; echo "== $cfg"; tail -6 $O/c51_layers_$tag.log | grep -E "conv_fwd|dinput"
env $cfg $T 200 python bench.py --gpus 1 --steps 40 --warmup 10 --no-roofline --no-cpu-baseline --no-fp32 > $O/c51_bench.json 2>>$O/c51_err.txt; python scratch/print_bench.py $O/c51_bench.json | cut -c1-200
done
