#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/c26_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c26_smoke.log
for cfg in "X=0" "MOPA_SCN_NO_BNBWD_FUSION=1" "MOPA_SCN_NO_BNBWD_FUSION=1 MOPA_TC_SPLIT=0"; do
  tag="${cfg// /_}"
  env $cfg $T 200 python tools/layer_table.py --out "$O/c26_layers_$tag.json" > "$O/c26_layers_$tag.log" 2>&1; echo "== $cfg"; tail -6 "$O/c26_layers_$tag.log" | grep -E "dinput|bn_bwd|conv_fwd"
  for rep in 1 2; do
  f="$O/c26_bench_${tag}_$rep.json"
  env $cfg $T 150 python bench.py --no-cpu-baseline --no-roofline --no-fp32 --steps 50 > "$f" 2>$O/c26_err.txt; echo "$cfg: $(python -c "import json,sys; d=json.load(open('$f')); print('%.3f ms/step  e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']))" 2>&1 | tail -1)"
  done
done
export MOPA_SCN_LIB=$PWD/scratch/bin/libmopa_scn_trace.so
for cfg in "X=0" "MOPA_SCN_NO_BNBWD_FUSION=1"; do
  echo "== trace $cfg"; env $cfg $T 120 python scratch/tc_trace3.py 2>&1 | tail -2
done
