#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
for v in 1 2 3 4 5 6; do
$T 200 python bench.py --steps 60 --warmup 5 --no-fp32 --no-roofline --no-cpu-baseline > $O/c40_bench_$v.json 2>>$O/c40_err.txt; echo -n "rep $v "; python scratch/print_bench.py $O/c40_bench_$v.json
done
$T 600 python -m pytest tests -x -q -m gpu > $O/c40_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c40_tests.log
