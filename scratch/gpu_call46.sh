#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
for cfg in "X=0" "MOPA_TC_SB=2"; do
tag="${cfg// /_}"
env $cfg $T 200 python tools/layer_table.py --out $O/c55_layers_$tag.json > $O/c55_layers_$tag.log 2>&1; echo "== $cfg"; tail -6 $O/c55_layers_$tag.log | grep -E "conv_fwd|dinput"
done
