#!/bin/bash
# usage: tools/regs.sh file.cu [flags]  -> registers / stack / spills per kernel (ptxas -v)
f=$(cd "$(dirname "$0")/../pseldnets_b200/csrc" && pwd)/$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -rdc=true "$@" -Xptxas -v -dc -o /tmp/regs_tmp.o $f 2>&1 | grep -E "Compiling entry|bytes stack|Used" | sed -e 's/ptxas info    ://' -e "s/Compiling entry function '_ZN4seld//" -e "s/' for 'sm_100a'//" | paste - - - | sed 's/Function properties for [^ ]*//' | cut -c1-230
