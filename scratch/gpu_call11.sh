#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
export MOPA_SCN_LIB=$PWD/scratch/bin/libmopa_scn_trace.so
for cfg in "X=0" "MOPA_TC_BN_RING=0" "MOPA_TC_DBG_NO_BNX=1" "MOPA_SCN_NO_BNSTATS_FUSION=1"; do
  echo "== $cfg"; env $cfg $T 120 python scratch/tc_trace3.py 2>&1 | tail -1
done
