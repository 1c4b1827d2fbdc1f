#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
for m in nvml smi none nvml smi none; do
BENCH_SAMPLER=$m $T 200 python bench.py --steps 20 --warmup 5 --no-fp32 --no-roofline --no-cpu-baseline > $O/c35_bench_$m.json 2>>$O/c35_err.txt; echo -n "$m "; python scratch/print_bench.py $O/c35_bench_$m.json
done
python -c "
import json; d=json.loads([l for l in open('$O/c35_bench_nvml.json') if l.startswith('{')][-1]); print(d['clocks'])"
tail -3 $O/c35_err.txt
