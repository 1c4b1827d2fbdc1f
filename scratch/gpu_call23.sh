#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 400 python tools/train_ab.py --steps 200 > $O/c34_train_ab.txt 2>$O/c34_err.txt; echo "rc=$?"; cat $O/c34_train_ab.txt; tail -3 $O/c34_err.txt
$T 200 python bench.py --steps 20 --warmup 5 --no-fp32 --no-roofline --no-cpu-baseline > $O/c34_bench_n1_20.json 2>>$O/c34_err.txt; python scratch/print_bench.py $O/c34_bench_n1_20.json
