#!/bin/bash
# Builds build/libegohmr_b200_trace.so: the product library with ONE source recompiled with a trace macro.
#   bash tools/build_trace_lib.sh gcn_umma_t EHB_K1_TRACE     (then tools/k1_trace.sh on the GPU box)
#   bash tools/build_trace_lib.sh conv_umma  EHB_CONV_TRACE   (then tools/conv_trace.sh "<launch ordinals>")
# Needs the objects of a normal build (python -c "import __graft_entry__ as g; g.build()").
set -e
SRC=${1:?source name without .cu}; MACRO=${2:?macro}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -D$MACRO -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -c egohmr_b200/csrc/$SRC.cu -o build/${SRC}_trace.o
OBJS=$(ls build/obj/*.o | grep -v "/$SRC.o")
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o build/libegohmr_b200_trace.so $OBJS build/${SRC}_trace.o
ls -la build/libegohmr_b200_trace.so
