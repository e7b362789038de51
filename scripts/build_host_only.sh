#!/bin/bash
# development: recompile only the host/C-ABI translation unit (qldpc_b200.cu + the headers it includes:
# OSD, SSF, frame sampler, utility kernels) and relink against the existing BP instantiation objects.
set -e
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC"
nvcc $F "$@" -c exp_ldpc_b200/csrc/qldpc_b200.cu -o build/full/qldpc_b200.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o exp_ldpc_b200/libqldpc_b200.so build/full/qldpc_b200.o build/full/bp_inst_*.o
