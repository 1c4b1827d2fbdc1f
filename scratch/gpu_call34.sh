#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02f_smoke.log
$T 400 python bench.py > $O/r02f_bench.json 2> $O/r02f_bench.err; python scratch/print_bench.py $O/r02f_bench.json
$T 200 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02f_bench_k20.json 2>> $O/r02f_bench.err; python scratch/print_bench.py $O/r02f_bench_k20.json
$T 300 python tools/layer_table.py --out $O/r02f_layers.json > $O/r02f_layers.log 2>&1; tail -6 $O/r02f_layers.log
$T 200 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/r02f_bench_reference.json 2>> $O/r02f_bench.err; cut -c1-400 $O/r02f_bench_reference.json
$T 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
