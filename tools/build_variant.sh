#!/bin/bash
# Developer A/B builds: a variant of libseldfeat.so with extra -D flags on ONE source file, the other objects cached.
#   tools/build_variant.sh out.so seld_foa_iv2.cu -DFOO -DBAR
set -e
out=$1; src=$2; shift 2
root="$(cd "$(dirname "$0")/.." && pwd)"; csrc=$root/pseldnets_b200/csrc; obj=$root/build/obj
mkdir -p $obj
FL="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -rdc=true -Xcompiler -fPIC"
objs=""
for f in seld_foa.cu seld_foa_iv2.cu seld_mic.cu seld_epilogue.cu seld_augment.cu seld_abi.cu; do
  if [ "$f" == "$src" ]; then
    o=$(mktemp --suffix=.o); nvcc $FL "$@" -dc -o $o $csrc/$f; objs="$objs $o"
  else
    o=$obj/${f%.cu}.o
    if [ ! -f $o ] || [ $csrc/$f -nt $o ] || [ $csrc/seld_plan.h -nt $o ] || [ $csrc/fft32.cuh -nt $o ] || [ $csrc/mel_seg.cuh -nt $o ]; then nvcc $FL -dc -o $o $csrc/$f; fi
    objs="$objs $o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o $out $objs -lcudadevrt
