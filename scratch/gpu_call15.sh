#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 400 python -m pytest tests/test_gpu_nccl.py -m gpu -q -x 2>&1 | tail -60 > $O/c15_nccl_test.log
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-fp32 > $O/c15_bench_n2.json 2> $O/c15_err.txt; echo "rc=$?"
tail -40 $O/c15_err.txt
tail -40 $O/c15_nccl_test.log
