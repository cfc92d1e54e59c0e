#!/bin/bash
# a compile-time variant of libiteres_gpu.so: tools/build_variant.sh <name> [-DMACRO=value ...] -> iteres_b200/csrc/variants/lib_<name>.so
# (A/B measurements only: ITX_LIB=<path> selects it in iteres_b200/capi.py)
set -e
name=$1; shift
cd "$(dirname "$0")/../iteres_b200/csrc"
mkdir -p variants
make -s itx_host.o itx_bgzf.o itx_bigwig.o itx_sam.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -c -o variants/itx_gpu_$name.o itx_gpu.cu 2> variants/ptxas_$name.log || { cat variants/ptxas_$name.log; exit 1; }
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_$name.so variants/itx_gpu_$name.o itx_host.o itx_bgzf.o itx_bigwig.o itx_sam.o -lz -lpthread -ldl
rm -f variants/itx_gpu_$name.o
grep -A2 "k_inflateILj4" variants/ptxas_$name.log | grep -E "Used" | sed "s/^/$name LG4: /"
