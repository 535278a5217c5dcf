#!/bin/bash
# build an experiment variant of libscipnp: tools/build_exp.sh <name> "<extra nvcc flags>"  (R = 4 warp-specialised instances only)
cd "$(dirname "$0")/../sci-algorithms_b200"
mkdir -p build/exp
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $2 -c csrc/ws_inst_r4.cu -o build/exp/ws_inst_r4_$1.o && \
nvcc -shared -gencode arch=compute_100a,code=sm_100a $(ls build/*.o | grep -v ws_inst_r4.o) build/exp/ws_inst_r4_$1.o -o build/exp/libscipnp_$1.so -cudart static && echo built $1
