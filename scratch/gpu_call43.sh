#!/bin/bash
O=gpurun_out; mkdir -p $O; TG=r02i
T="timeout -k 5"
$T 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TG}_smoke.log
$T 900 python -m pytest tests -q -m gpu 2>&1 | tail -2
$T 400 python bench.py > $O/${TG}_bench.json 2> $O/${TG}_bench.err; python scratch/print_bench.py $O/${TG}_bench.json
$T 200 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TG}_bench_k20.json 2>> $O/${TG}_bench.err; python scratch/print_bench.py $O/${TG}_bench_k20.json
$T 300 python tools/layer_table.py --out $O/${TG}_layers.json > $O/${TG}_layers.log 2>&1; tail -6 $O/${TG}_layers.log
