#!/bin/bash
# timing-only ablations of the forward kernel on zero inputs (SM clock at max): usage gpu_ablate.sh v1 v2 ...
mkdir -p gpurun_out; L=gpurun_out/ablate.log; : > $L
T=tools/fa_selftest
for v in "$@"; do
  export LD_LIBRARY_PATH=$PWD/build/$v
  for args in "4 32 8192 128 1 0 0 Z 10" "4 32 8192 64 1 0 0 Z 10"; do
    echo "##### $v: $args" >> $L
    timeout 200 $T attn $args 2>&1 | grep -E "TIMING|watchdog|error" >> $L
  done
done
cat $L | cut -c1-200
timeout 300 python tools/cudnn_zero.py 2>&1 | grep -E "ZERO_VS|Error|error" | tee -a $L
