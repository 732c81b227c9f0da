#!/bin/bash
# Developer helper: build libtfhe_b200.so with extra -D flags for br_cggi32.cu into tfhe_gpu_b200/build/ab/NAME.so
# (kernel A/B measurements with tools/abbench.cpp).  usage: tools/build_variant.sh NAME [-DFLAG ...]
set -e
cd "$(dirname "$0")/../tfhe_gpu_b200"
name=$1; shift
mkdir -p build/ab
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -O3 -Xcompiler -fopenmp --expt-relaxed-constexpr -Xptxas -v"
$NV "$@" -c csrc/br_cggi32.cu -o build/ab/br_cggi32_$name.o 2> build/ab/br_cggi32_$name.ptxas.log || (tail -20 build/ab/br_cggi32_$name.ptxas.log; false)
objs=$(ls build/*.o | grep -v br_cggi32.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/ab/$name.so $objs build/ab/br_cggi32_$name.o -lcudart -lgomp -ldl
grep -A3 "ILi10ELi4ELi4ELb1ELb0ELb0ELi0ELb.EEEvNS_10CGGI32ArgsE' for" build/ab/br_cggi32_$name.ptxas.log | grep -E "Used|spill stores" | tr '\n' ' '; echo
