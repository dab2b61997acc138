#!/bin/bash
# tuning helper: builds libbridge_b200 variants that differ only in bb_theta.cu macros -> bridge.jl_b200/lib/var/<name>.so
# usage: tools/build_theta_variants.sh name1="-DBB_TDEPTH=3" name2="-DBB_TWPF=0 -DBB_TDEPTH=4" ...
set -e
cd "$(dirname "$0")/../bridge.jl_b200"
make -s -j8 -C csrc
mkdir -p lib/var build/var
OBJS=$(ls build/*.o | grep -v bb_theta.o)
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 \
       $flags -Xptxas -v -c csrc/bb_theta.cu -o build/var/bb_theta_$name.o 2>&1 | grep -E "forward_kernelI8MFhnHypoLi1|registers|spill" | grep -A2 "forward_kernelI8MFhnHypoLi1" | grep -E "registers|spill" | tr '\n' ' '
  echo " <- $name ($flags)"
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/var/$name.so $OBJS build/var/bb_theta_$name.o
done
