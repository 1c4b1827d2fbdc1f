#!/bin/bash
# debug build of the library with the device-side timeline of k_conv_tc (MOPA_TC_TRACE) -> scratch/bin/libmopa_scn_trace.so
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch/bin /tmp/mopa_trace
for f in geometry conv conv_tc conv_dw_tc bn_io program xm_ops vgi; do
  nvcc -DMOPA_TC_TRACE -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I mopa_b200/csrc -c mopa_b200/csrc/$f.cu -o /tmp/mopa_trace/$f.o &
done
wait
nvcc -shared -o scratch/bin/libmopa_scn_trace.so /tmp/mopa_trace/*.o -lcudart_static -lpthread -ldl -lrt
ls -la scratch/bin/libmopa_scn_trace.so
