#!/bin/bash
# round-2 call 3: tile-rulebook conv kernel with one gather warp per stage; every command under a hard kill
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/c3_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/c3_smoke.log
if ! grep -q "^smoke:" $O/c3_smoke.log; then echo "SMOKE FAILED - stopping"; exit 1; fi
$T 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 > $O/c3_tests.log; tail -4 $O/c3_tests.log
$T 200 python bench.py --no-cpu-baseline > $O/c3_bench.json 2> $O/c3_bench.err; cut -c1-300 $O/c3_bench.json; tail -3 $O/c3_bench.err
$T 200 python tools/layer_table.py --out $O/c3_layers.json > $O/c3_layers.log 2>&1; tail -8 $O/c3_layers.log
for cfg in "MOPA_TC_SA=4" "MOPA_TC_CTAS=1" "MOPA_TC_CTAS=1 MOPA_TC_SA=8" "MOPA_TC_TPC=1" "MOPA_TC_TPC=1 MOPA_TC_SA=4" "MOPA_TC_TPC=2"; do
  env $cfg $T 120 python bench.py --no-cpu-baseline --no-roofline --steps 20 > "$O/c3_bench_${cfg// /_}.json" 2>/dev/null; echo "$cfg: $(cut -c60-200 "$O/c3_bench_${cfg// /_}.json")"
done
$T 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py 2>&1 | tail -15 > $O/c3_tests_new.log; tail -8 $O/c3_tests_new.log
