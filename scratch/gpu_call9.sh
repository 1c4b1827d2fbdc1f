#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
for cfg in "X=0" "MOPA_TC_SPLIT=0 MOPA_SCN_NO_BNSTATS_FUSION=1" "MOPA_TC_DBG_NO_BNX=1" "MOPA_SCN_NO_BNSTATS_FUSION=1"; do
  tag="${cfg// /_}"
  env $cfg $T 200 python tools/layer_table.py --out "$O/c21_layers_$tag.json" > "$O/c21_layers_$tag.log" 2>&1; echo "== $cfg"; tail -6 "$O/c21_layers_$tag.log"
done
$T 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_mopa_step.py -x -q -m gpu > $O/c21_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/c21_tests.log
