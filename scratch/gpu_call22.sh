#!/bin/bash
T="timeout -k 5"
for m in default reserve noar; do $T 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scratch/slow_steps.py $m 2>&1 | grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" | tail -3; done
