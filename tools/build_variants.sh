#!/bin/bash
# development aid: build differently tuned copies of the library into build/variants/<name>.so
# usage: tools/build_variants.sh name1 "-DFLAG=.. -DFLAG2=.." name2 "..." ...
set -e
cd "$(dirname "$0")/../tauray_b200/csrc"
mkdir -p ../../build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false --expt-extended-lambda -Xcompiler -fPIC"
while [ $# -gt 1 ]; do
  name=$1; extra=$2; shift 2
  ( nvcc $FLAGS $extra -Xptxas -v -c tr_render.cu -o ../../build/variants/$name.o 2> ../../build/variants/$name.ptxas.log
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/variants/$name.so tr_bvh.o ../../build/variants/$name.o tr_abi.o -lcudart_static -lpthread -ldl -lrt
    echo "built $name: $(grep -A3 'Compiling.*k_wf_trace\|Compiling.*k_wf_shade' ../../build/variants/$name.ptxas.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')" ) &
done
wait
