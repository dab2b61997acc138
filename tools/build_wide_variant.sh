#!/bin/bash
# tuning helper: build bridge.jl_b200/lib/var/<name>.so with extra -D flags applied to the translation units of the
# d' >= 2 models (their kernels and launch code are self-contained per unit)
# usage: tools/build_wide_variant.sh <name> "<flags>"
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../bridge.jl_b200/csrc"
mkdir -p ../lib/var ../build/var
units="bb_inst_linpro3 bb_inst_fhn_diag bb_inst_linpro2 bb_inst_lorenz"
objs=""; excl=""
for u in $units; do
  obj=../build/var/${name}_$u.o
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -I../build -Xcompiler -fPIC,-ffp-contract=off,-O2 $flags -Xptxas -v -c $u.cu -o $obj 2> ../build/var/${name}_$u.ptxas &
  objs="$objs $obj"; excl="$excl -e /$u.o"
done
wait
others=$(ls ../build/*.o | grep -v $excl)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/var/$name.so $objs $others -ldl
echo built ../lib/var/$name.so
