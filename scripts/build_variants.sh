#!/bin/bash
# Developer probe: build A/B variants of libnerf_b200.so that differ only in composite.cu tuning macros.
# usage: scripts/build_variants.sh name "-DNB_FWD_R64=2 -DNB_FWD_MB64=6" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/../nerf_simple_b200/csrc"
make -s >/dev/null 2>&1
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c composite.cu -o build/variants/composite_$name.o
  objs=$(ls build/*.o | grep -v composite.o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so $objs build/variants/composite_$name.o -cudart static
  echo built build/variants/lib_$name.so
done
