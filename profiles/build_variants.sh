#!/bin/bash
# Build A/B variants of libb200seg.so with different -D knobs into gpurun_out/variants/ (scratch; selected at run time
# with B200SEG_LIB=<path>).   usage: profiles/build_variants.sh name "-DK4_MIN_CTAS=4 -DK4_PAIR_ROWS=0" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
CS=rnd_semantic_segmentation_b200/csrc
mkdir -p variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  d=/tmp/b200seg_var_$name; mkdir -p $d
  for f in api eval_kernels ce_kernels softce_kernels gemm_sm100 aspp_head; do
    ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
        --expt-relaxed-constexpr $flags -c $CS/$f.cu -o $d/$f.o ) &
  done
  wait
  nvcc -shared -o variants/libb200seg_$name.so $d/*.o -gencode arch=compute_100a,code=sm_100a -cudart static -Xcompiler -fPIC
  echo built variants/libb200seg_$name.so
done
