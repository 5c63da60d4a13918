#!/bin/bash
# Instrumented variant of the library (in-kernel phase clocks of attn_fused_kernel): libmst_b200_prof.so next to the real one.
#   MST_LIB_PATH=mastermetastyletransfer_b200/libmst_b200_prof.so python tools/attn_prof.py
set -e
cd "$(dirname "$0")/.."
python -m mastermetastyletransfer_b200.csrc.build > /dev/null
OBJ=mastermetastyletransfer_b200/csrc/_obj
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DMST_AF_PROF -c mastermetastyletransfer_b200/csrc/attn_fused.cu -o /tmp/attn_fused_prof.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DMST_MLP_PROF -c mastermetastyletransfer_b200/csrc/mlp_fused.cu -o /tmp/mlp_fused_prof.o
nvcc -shared -o mastermetastyletransfer_b200/libmst_b200_prof.so $(ls $OBJ/*.o | grep -v -e attn_fused.o -e mlp_fused.o) /tmp/attn_fused_prof.o /tmp/mlp_fused_prof.o -gencode arch=compute_100a,code=sm_100a
echo built mastermetastyletransfer_b200/libmst_b200_prof.so
