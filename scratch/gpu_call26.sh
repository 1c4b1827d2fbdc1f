#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
for v in 0 1000000 2000000 4000000 8000000 0; do
MOPA_SCN_BN_FUSED_MAX=$v $T 200 python bench.py --steps 40 --warmup 10 --no-fp32 --no-roofline --no-cpu-baseline > $O/c38_bench_$v.json 2>>$O/c38_err.txt; echo -n "max=$v "; python scratch/print_bench.py $O/c38_bench_$v.json
done
