#!/bin/bash
# usage: bash tools/build_variant.sh NAME "-DMACRO=VALUE ..."   -> bench_kernels/var_NAME.so
# A/B builds of the whole library with extra compile-time macros (tile shapes, occupancy targets);
# run one with BARCODER_B200_LIB=bench_kernels/var_NAME.so python bench.py ...  (tools/run_variants.sh)
set -e
cd "$(dirname "$0")/../barcoder_b200/csrc"
name=$1; shift
objs=""
mkdir -p /tmp/bcvar_$name
for f in bc_api bc_kernels bc_join bc_cjoin bc_sort bc_guides; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $f.cu -o /tmp/bcvar_$name/$f.o &
  objs="$objs /tmp/bcvar_$name/$f.o"
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../bench_kernels/var_$name.so $objs -lcudart
echo built bench_kernels/var_$name.so
