#!/bin/bash
# A/B builds of one translation unit (default: the sparse forward kernel; TU=markowitz_bwd_sparse_f32 for the backward):
#   tools/ab_variants.sh name1:"-DFLAG ..." name2:"..." -> deepdow_b200/variants/libddw_<name>.so
# (every other object comes from the product build in deepdow_b200/build).  Time them with tools/time_fwd.py.
set -e
cd "$(dirname "$0")/.."
TU=${TU:-markowitz_fwd_sparse_f32}
mkdir -p deepdow_b200/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I deepdow_b200/csrc $flags \
       -c deepdow_b200/csrc/$TU.cu -o /tmp/ab_$name.o
  objs=$(ls deepdow_b200/build/*.o | grep -v "/$TU.o")
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o deepdow_b200/variants/libddw_$name.so $objs /tmp/ab_$name.o
  echo built $name "($flags)"
done
