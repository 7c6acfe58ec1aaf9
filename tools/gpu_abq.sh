#!/bin/bash
# item-boundary experiments (zero inputs; qc / qd give wrong results by construction): usage gpu_abq.sh v1 v2 ...
mkdir -p gpurun_out; L=gpurun_out/abq.log; : > $L
T=tools/fa_selftest
for r in 1 2; do
for v in "$@"; do
  export LD_LIBRARY_PATH=$PWD/build/$v
  for args in "8 16 1024 64 0 0 0 Z 40" "4 32 8192 64 1 0 0 Z 10" "4 32 8192 128 1 1 0 Z 10" "4 32 8192 128 1 0 0 Z 10"; do
    timeout 100 $T attn $args 2>&1 | grep -E "TIMING|watchdog|error" | tee -a $L | cut -c1-200 | sed "s/^/$v: /" | sed 's/TIMING attn //; s/median //; s/ -> / /'
  done
done
done
