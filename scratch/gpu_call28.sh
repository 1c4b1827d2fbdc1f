#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
for v in 1 2 3 4 5 6 7 8; do
$T 200 python bench.py --steps 60 --warmup 5 --no-fp32 --no-roofline --no-cpu-baseline > $O/c43_bench_$v.json 2>>$O/c43_err.txt; echo -n "$v "; python scratch/print_bench.py $O/c43_bench_$v.json | cut -c1-250
done
