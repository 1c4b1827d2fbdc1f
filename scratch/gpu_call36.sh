#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
$T 900 python tools/sweep.py --out $O/r02g_sweep.json > $O/r02g_sweep.log 2>&1; cat $O/r02g_sweep.log
$T 400 python bench.py --steps 20 --warmup 5 > $O/r02g_bench_k20.json 2>$O/r02g_err.txt; python scratch/print_bench.py $O/r02g_bench_k20.json; python -c "
import json; d=json.loads([l for l in open('$O/r02g_bench_k20.json') if l.startswith('{')][-1]); print(d['roofline']['whole_step'], d['roofline']['frac'])"
