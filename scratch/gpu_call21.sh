#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
for kw in "20 5" "40 10"; do set -- $kw
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps $1 --warmup $2 --no-fp32 --no-roofline > $O/c32_bench_n2_$1.json 2>$O/c32_err.txt; echo "rc=$?"
python scratch/print_bench.py $O/c32_bench_n2_$1.json
done
$T 200 python bench.py --steps 20 --warmup 5 --no-fp32 --no-roofline --no-cpu-baseline > $O/c32_bench_n1_20.json 2>>$O/c32_err.txt; python scratch/print_bench.py $O/c32_bench_n1_20.json
