#!/bin/bash
# Developer build with -DSIHL_PHASE_TIMING into sihl_b200/lib/libsihl_b200_dbg.so (not used by the product)
set -e
cd "$(dirname "$0")/../sihl_b200/csrc"
mkdir -p /tmp/sihl_dbg
for f in od_api od_anchors od_assign od_quad od_loss od_exchange od_infer od_nms; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -DSIHL_PHASE_TIMING -Xcompiler -fPIC -c $f.cu -o /tmp/sihl_dbg/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -o ../lib/libsihl_b200_dbg.so /tmp/sihl_dbg/od_*.o -cudart shared
