#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/c46_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c46_smoke.log
if ! grep -q "^smoke:" $O/c46_smoke.log; then echo "SMOKE FAILED - stopping"; tail -30 $O/c46_smoke.log; exit 1; fi
$T 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k tail > $O/c46_tail.log 2>&1; echo "tail test rc=$?"; tail -3 $O/c46_tail.log
for cfg in "X=0" "MOPA_TC_TAIL=0" "X=0" "MOPA_TC_TAIL=0"; do
env $cfg $T 200 python bench.py --gpus 1 --steps 40 --warmup 10 --no-roofline --no-cpu-baseline --no-fp32 > $O/c46_bench.json 2>>$O/c46_err.txt; echo -n "$cfg "; python scratch/print_bench.py $O/c46_bench.json | cut -c1-200
done
for cfg in "X=0" "MOPA_TC_TAIL=0"; do
env $cfg $T 200 python tools/layer_table.py --out $O/c46_layers_$cfg.json > $O/c46_layers_$cfg.log 2>&1; echo "== $cfg"; tail -6 $O/c46_layers_$cfg.log | grep -E "conv_fwd|dinput"
done
$T 900 python -m pytest tests -x -q -m gpu > $O/c46_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c46_tests.log
