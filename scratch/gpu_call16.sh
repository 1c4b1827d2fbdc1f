#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 400 python -m pytest tests/test_gpu_nccl.py -m gpu -q -x 2>&1 | tail -5
$T 300 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --no-fp32 > $O/c16_n2.log 2>&1; echo "rc=$?"
grep -v "^\*\*\*\|OMP_NUM" $O/c16_n2.log | cut -c1-600 | tail -20
