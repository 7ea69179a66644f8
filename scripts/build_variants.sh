#!/bin/bash
# Developer probe: build A/B variants of libnerf_b200.so that differ only in the tuning macros of one source.
# usage: scripts/build_variants.sh <source.cu> name "-DFOO=2 -DBAR=6" [name2 "flags2" ...]
# then:  NERF_B200_LIB=nerf_simple_b200/csrc/build/variants/lib_<name>.so python scripts/...
set -e
cd "$(dirname "$0")/../nerf_simple_b200/csrc"
src=$1; shift
stem=${src%.cu}
make -s >/dev/null 2>&1
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c $src -o build/variants/${stem}_$name.o
  objs=$(ls build/*.o | grep -v "build/$stem.o")
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so $objs build/variants/${stem}_$name.o -cudart static
  echo built build/variants/lib_$name.so
done
