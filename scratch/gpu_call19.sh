#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
for rep in 1 2 3; do
  f="$O/c30_bench_$rep.json"
  $T 200 python bench.py --no-cpu-baseline --no-roofline --no-fp32 --steps 200 > "$f" 2>$O/c30_err.txt; echo "rep $rep: $(python -c "import json,sys; d=json.load(open('$f')); print('%.3f ms/step median %.3f max %.2f slow %s | e2e %.3f median %.3f' % (d['ms_per_step'], d['step_ms']['median'], d['step_ms']['max'], d['step_ms']['slow_steps_rank0'], d['e2e']['ms_per_step'], d['e2e']['median_ms']))" 2>&1 | tail -1)"
done
