#!/bin/bash
# tuning helper: build bridge.jl_b200/lib/var/<name>.so with extra -D flags applied to ONE translation unit
# usage: tools/build_variant.sh <name> <unit.cu> "<flags>"      (the other objects come from the regular build)
set -e
name=$1; unit=$2; flags=$3
cd "$(dirname "$0")/../bridge.jl_b200/csrc"
mkdir -p ../lib/var ../build/var
obj=../build/var/${name}_$(basename $unit .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 $flags -c $unit -o $obj
others=$(ls ../build/*.o | grep -v "/$(basename $unit .cu).o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/var/$name.so $obj $others -ldl
echo built ../lib/var/$name.so
