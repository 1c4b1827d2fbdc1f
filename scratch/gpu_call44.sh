#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/c54_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c54_smoke.log
if ! grep -q "^smoke:" $O/c54_smoke.log; then echo "SMOKE FAILED - stopping"; tail -30 $O/c54_smoke.log; exit 1; fi
$T 600 python -m pytest tests -x -q -m gpu > $O/c54_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c54_tests.log
for v in 1 0; do MOPA_SCN_ARENA_POOL=$v $T 120 python scratch/empty_cache_loop.py 2>&1 | tail -1; done
for v in 1 0; do
MOPA_SCN_ARENA_POOL=$v $T 200 python bench.py --gpus 1 --steps 40 --warmup 10 --no-roofline --no-cpu-baseline --no-fp32 > $O/c54_bench_$v.json 2>>$O/c54_err.txt; echo -n "pool=$v "; python scratch/print_bench.py $O/c54_bench_$v.json | cut -c1-200
done
