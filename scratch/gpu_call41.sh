#!/bin/bash
O=gpurun_out; TG=r02h; mkdir -p $O
T="timeout -k 5"
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-fp32"
$T 300 ncu --set full --clock-control none --import-source on -k regex:k_dw_rows_mma -s 4 -c 1 -o /tmp/${TG}_rows $B > /dev/null 2>&1
ncu -i /tmp/${TG}_rows.ncu-rep --page raw --csv > $O/${TG}_rows_raw.csv 2>/dev/null
ncu -i /tmp/${TG}_rows.ncu-rep --page source --csv > $O/${TG}_rows_src1.csv 2>/dev/null
ls -la $O/${TG}_rows*
