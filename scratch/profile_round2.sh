#!/bin/bash
# One gpurun call that produces the round-2 artefacts (tag = $1): every command under a hard kill.
O=gpurun_out; TG=${1:-r02}; mkdir -p $O
T="timeout -k 5"
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/${TG}_smoke.log
if ! grep -q "xm operators ok" $O/${TG}_smoke.log; then echo "SMOKE FAILED - stopping"; exit 1; fi
$T 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/${TG}_tests.log; tail -6 $O/${TG}_tests.log
$T 300 python tools/grad_error_table.py --sensor nuscenes --out $O/${TG}_grad_errors > $O/${TG}_grad_errors.log 2>&1; tail -5 $O/${TG}_grad_errors.log
$T 300 python tools/grad_error_table.py --sensor kitti --out $O/${TG}_grad_errors_kitti > /dev/null 2>&1
$T 400 python bench.py > $O/${TG}_bench.json 2> $O/${TG}_bench.err; python scratch/print_bench.py $O/${TG}_bench.json; tail -2 $O/${TG}_bench.err
$T 200 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TG}_bench_k20.json 2>> $O/${TG}_bench.err; python scratch/print_bench.py $O/${TG}_bench_k20.json
$T 200 python tools/train_ab.py --steps 200 > $O/${TG}_train_ab.txt 2>> $O/${TG}_bench.err; tail -8 $O/${TG}_train_ab.txt
$T 300 python tools/layer_table.py --out $O/${TG}_layers.json > $O/${TG}_layers.log 2>&1; tail -7 $O/${TG}_layers.log
$T 900 python tools/sweep.py --out $O/${TG}_sweep.json > $O/${TG}_sweep.log 2>&1; cat $O/${TG}_sweep.log
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-fp32"
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TG}_launches.csv $B > /dev/null 2>&1
export MOPA_SCN_NO_DW_OVERLAP=1
$T 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_conv_tc -c 260 --csv --log-file $O/${TG}_conv_tc_traffic.csv $B > /dev/null 2>&1
for spec in "k_conv_tc 150 4 conv_tc" "k_dw_tc\$ 78 3 dw_tc" "k_bn_stats 100 3 bn_stats"; do
  set -- $spec
  $T 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o /tmp/${TG}_$4 $B > /dev/null 2>&1
  ncu -i /tmp/${TG}_$4.ncu-rep --page raw --csv > $O/${TG}_$4_raw.csv 2>/dev/null
  ncu -i /tmp/${TG}_$4.ncu-rep --page source --csv --kernel-id ::regex:$1:1 > $O/${TG}_$4_src1.csv 2>/dev/null
done
ls -la $O | grep ${TG}_ ; du -sh $O
