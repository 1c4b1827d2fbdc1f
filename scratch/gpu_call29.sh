#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 900 python tools/sweep.py --out $O/r02e_sweep.json > $O/r02e_sweep.log 2>&1; cat $O/r02e_sweep.log
for v in 1 2 3; do
$T 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-roofline --no-cpu-baseline --no-fp32 > $O/c44_bench_$v.json 2>>$O/c44_err.txt; python scratch/print_bench.py $O/c44_bench_$v.json | cut -c1-230
done
