#!/bin/bash
# One gpurun call: GPU tests, bench line, per-layer table, ncu launch list, ncu traffic of every k_conv_tc launch of one
# step, ncu --set full of a few launches of the three main kernels (exported to CSV on the box; reports stay small).
O=gpurun_out; T=${1:-r6}
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/${T}_tests.log; tail -2 $O/${T}_tests.log
timeout 400 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; cut -c1-160 $O/${T}_bench.json
timeout 300 python tools/layer_table.py --out $O/${T}_layers.json > $O/${T}_layers.log 2>&1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 825 -c 300 --csv --log-file $O/${T}_launches.csv $B > /dev/null 2>&1
export MOPA_SCN_NO_DW_OVERLAP=1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_conv_tc -s 147 -c 49 --csv --log-file $O/${T}_conv_tc_traffic.csv $B > /dev/null 2>&1
for spec in "k_conv_tc 147 4 conv_tc" "k_dw_tc\$ 78 4 dw_tc" "k_bn_fused 156 4 bn_fused"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o /tmp/${T}_$4 $B > /dev/null 2>&1
  ncu -i /tmp/${T}_$4.ncu-rep --page raw --csv > $O/${T}_$4_raw.csv 2>/dev/null
  ncu -i /tmp/${T}_$4.ncu-rep --page source --csv --kernel-id ::regex:$1:1 > $O/${T}_$4_src1.csv 2>/dev/null
  ncu -i /tmp/${T}_$4.ncu-rep --page source --csv --kernel-id ::regex:$1:2 > $O/${T}_$4_src2.csv 2>/dev/null
done
ls -la $O | grep ${T}_ ; du -sh $O
