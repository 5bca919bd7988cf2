#!/bin/bash
# Build tuning variants of libkatgpu.so (different -D shapes of the sweep kernel) into
# kat_b200/lib/variants/ — they travel to the GPU box with gpurun and are selected with KATGPU_LIB.
#   usage: tools_variants.sh name "-DKATGPU_SWEEP_UNROLL=2 -DKATGPU_SWEEP_MIN_CTAS=4" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p kat_b200/lib/variants build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  d=build/variants/$name; mkdir -p $d
  for f in count_kernels partition_kernels tile_kernels dist_kernels parse_kernels reduce_kernels katgpu_api katgpu_fastq katgpu_walks katgpu_dist; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $flags -c kat_b200/csrc/$f.cu -o $d/$f.o &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o kat_b200/lib/variants/libkatgpu_$name.so $d/*.o -ldl
  echo "built kat_b200/lib/variants/libkatgpu_$name.so ($flags)"
done
