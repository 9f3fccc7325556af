#!/bin/bash
# tools/build_alt.sh NAME "-DFLAG ..." : alternative libhmmcuda build (ring_viterbi.cu recompiled with the flags) into
# tools/alt/libhmmcuda_NAME.so; run anything with LIBHMMCUDA=$PWD/tools/alt/libhmmcuda_NAME.so
set -e
cd "$(dirname "$0")/../hmmspikesorter.jl_b200/csrc"
make -s -j8
mkdir -p ../../tools/alt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $2 -c ring_viterbi.cu -o /tmp/alt_rv_$1.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/alt/libhmmcuda_$1.so api.o model.o faithful.o generic_parallel.o reconstruct.o /tmp/alt_rv_$1.o ring_em.o update.o -lcudart -lpthread
echo built tools/alt/libhmmcuda_$1.so
