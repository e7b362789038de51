#!/bin/bash
# development build: only the kernel variants the benchmark codes need (-DQD_FAST_BUILD)
set -e
cd /root/repo/exp_ldpc_b200/csrc
nvcc -DQD_FAST_BUILD -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --extended-lambda \
     -Xcompiler -fPIC -shared -Xptxas -v -o ../libqldpc_b200.so qldpc_b200.cu "$@"
