#!/bin/bash
# timing-only ablations at d = 64 on zero inputs (HEAD protocol): usage gpu_ablate64.sh v1 v2 ...
mkdir -p gpurun_out; L=gpurun_out/ablate64.log; : > $L
T=tools/fa_selftest
for v in "$@"; do
  export LD_LIBRARY_PATH=$PWD/build/$v
  for args in "4 32 8192 64 1 0 0 Z 10" "8 16 1024 64 0 0 0 Z 30"; do
    timeout 200 $T attn $args 2>&1 | grep -E "TIMING" | sed "s/^/$v: /" | sed 's/TIMING attn //' >> $L
  done
done
cat $L | cut -c1-160
