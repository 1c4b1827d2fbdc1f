#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-fp32 > $O/r02j_bench_n2.json 2>$O/r02j_err.txt; echo "rc=$?"
python scratch/print_bench.py $O/r02j_bench_n2.json
$T 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-fp32 --no-roofline --no-cpu-baseline > $O/r02j_bench_n1.json 2>>$O/r02j_err.txt; python scratch/print_bench.py $O/r02j_bench_n1.json
$T 200 python bench.py --sensor kitti --steps 30 --warmup 10 --no-fp32 --no-roofline --no-cpu-baseline > $O/r02j_bench_kitti.json 2>>$O/r02j_err.txt; python scratch/print_bench.py $O/r02j_bench_kitti.json
