#!/bin/bash
# usage: scripts/ptxas_v.sh <file.cu> [extra nvcc flags]: registers / spills of every kernel in one source file
cd "$(dirname "$0")/.." || exit 1
f=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v "$@" \
  -c "lightweight-human-pose-estimation.pytorch_b200/csrc/$f" -o /tmp/ptxas_v.o 2>&1 | grep -i "error\|spill\|Used\|Compiling entry" | sed 's/ptxas info    : //'
