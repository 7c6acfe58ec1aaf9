#!/bin/bash
# d <= 64 shapes, zero and Set-S inputs, variants in turn: usage gpu_ab64.sh ROUNDS v1 v2 ...
R=$1; shift
mkdir -p gpurun_out; L=gpurun_out/ab64.log; : > $L
T=tools/fa_selftest
for r in $(seq 1 $R); do
for v in "$@"; do
  export LD_LIBRARY_PATH=$PWD/build/$v
  for args in "4 32 8192 64 1 0 0 Z 10" "4 32 8192 64 1 0 0 S 20" "4 32 8192 64 1 1 0 S 20" "8 16 1024 64 0 0 0 S 40" "8 16 1024 64 0 0 0 Z 40" "2 16 4096 32 1 0 0 S 20"; do
    timeout 200 $T attn $args 2>&1 | grep -E "TIMING|FAIL|watchdog|error" | tee -a $L | cut -c1-200 | sed "s/^/$v: /" | sed 's/TIMING attn //; s/median //; s/ -> / /'
  done
done
done
