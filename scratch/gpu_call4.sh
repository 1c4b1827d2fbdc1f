#!/bin/bash
# round-2 call 4: device timelines of the tile-rulebook conv kernel + xm operator tests; every command under a hard kill
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 300 python -m pytest tests/test_gpu_xm.py -m gpu -x -q 2>&1 | tail -15 > $O/c4_tests_xm.log; tail -12 $O/c4_tests_xm.log
export MOPA_SCN_LIB=$PWD/scratch/bin/libmopa_scn_trace.so
for cfg in "" "MOPA_TC_SA=4" "MOPA_TC_CTAS=1 MOPA_TC_SA=8"; do
  echo "#### config: $cfg" >> $O/c4_trace.txt
  env $cfg $T 200 python scratch/tc_trace2.py 0 16 16 1 32 32 1 64 32 3 64 64 4 160 80 >> $O/c4_trace.txt 2>&1
done
cat $O/c4_trace.txt
