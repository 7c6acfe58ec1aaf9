#!/bin/bash
# zero-input timing (SM clock stays at max): cycle-domain comparison of library variants
mkdir -p gpurun_out; L=gpurun_out/trip5_zero.log; : > $L
T=tools/fa_selftest
for r in 1 2; do
for v in "$@"; do
  export LD_LIBRARY_PATH=$PWD/build/$v
  for args in "4 32 8192 128 1 0 0 Z 20" "4 32 8192 128 1 0 0 S 20" "4 32 8192 64 1 0 0 Z 20" "4 32 8192 64 1 0 0 S 20" "8 16 1024 64 0 0 0 Z 30"; do
    echo "##### $v: $args" >> $L
    timeout 200 $T attn $args >> $L 2>&1; echo "exit=$?" >> $L
  done
done
done
grep -E "#####|FAIL|TIMING|exit=[1-9]" $L | cut -c1-200
