#!/bin/bash
O=gpurun_out; TG=r02h; mkdir -p $O
T="timeout -k 5"
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-fp32"
export MOPA_SCN_NO_DW_OVERLAP=1
$T 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_conv_tc -c 260 --csv --log-file $O/${TG}_conv_tc_traffic.csv $B > /dev/null 2>&1
for spec in "k_conv_tc 150 4 conv_tc" "k_dw_tc\$ 78 3 dw_tc" "k_bn_stats 100 3 bn_stats"; do
  set -- $spec
  $T 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o /tmp/${TG}_$4 $B > /dev/null 2>&1
  ncu -i /tmp/${TG}_$4.ncu-rep --page raw --csv > $O/${TG}_$4_raw.csv 2>/dev/null
  ncu -i /tmp/${TG}_$4.ncu-rep --page source --csv --kernel-id ::regex:$1:1 > $O/${TG}_$4_src1.csv 2>/dev/null
done
ls -la $O | grep ${TG}_ | head -20
