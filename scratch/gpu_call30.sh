#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/c45_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c45_smoke.log
if ! grep -q "^smoke:" $O/c45_smoke.log; then echo "SMOKE FAILED - stopping"; tail -30 $O/c45_smoke.log; exit 1; fi
for v in 1 2 3; do
$T 200 python bench.py --gpus 1 --steps 40 --warmup 10 --no-roofline --no-cpu-baseline --no-fp32 > $O/c45_bench_$v.json 2>>$O/c45_err.txt; python scratch/print_bench.py $O/c45_bench_$v.json | cut -c1-230
done
$T 200 python bench.py --sensor kitti --steps 30 --warmup 10 --no-roofline --no-cpu-baseline --no-fp32 > $O/c45_bench_kitti.json 2>>$O/c45_err.txt; python scratch/print_bench.py $O/c45_bench_kitti.json | cut -c1-230
$T 900 python -m pytest tests -x -q -m gpu > $O/c45_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c45_tests.log
