#!/bin/bash
# 2 GPUs: NCCL correctness test + the data-parallel bench at N=2 (torchrun, as the driver launches it)
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
nvidia-smi -L | head -4
$T 400 python -m pytest tests/test_gpu_nccl.py -m gpu -q 2>&1 | tail -4
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 10 --no-fp32 > $O/c14_bench_n2.json 2> $O/c14_err.txt; python -c "
import json; d=json.load(open('$O/c14_bench_n2.json')); print('N=2: %.3f ms/step  %.1f M pts/s  all-reduce median %.0f us  per-rank %s imbalance %s' % (d['ms_per_step'], d['value']/1e6, d['all_reduce_us_median'], [round(r['ms_per_step'],3) for r in d['per_rank']], d['imbalance']))" 2>&1 | tail -2
$T 200 python bench.py --steps 50 --warmup 10 --no-fp32 --no-cpu-baseline --no-roofline > $O/c14_bench_n1.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/c14_bench_n1.json')); print('N=1: %.3f ms/step  %.1f M pts/s' % (d['ms_per_step'], d['value']/1e6))"
$T 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-300
