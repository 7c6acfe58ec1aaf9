#!/bin/bash
# timing-only sweep over library variants: fa_selftest attn cases given after "--"
mkdir -p gpurun_out
L=gpurun_out/sweep2.log
: > $L
T=tools/fa_selftest
for v in "$@"; do
  echo "##### variant $v" >> $L
  export LD_LIBRARY_PATH=$PWD/build/$v
  for args in "4 32 8192 128 1 0 0 S 30" "4 32 8192 128 1 1 0 S 30" "8 16 1024 64 0 0 0 S 30" "2 16 4096 64 1 1 0 S 30" "1 32 16384 128 1 1 0 S 10" "16 16 2048 128 1 1 0 S 30"; do
    timeout 200 $T attn $args >> $L 2>&1; echo "exit=$?" >> $L
  done
done
grep -E "#####|FAIL|TIMING|exit=[1-9]" $L | cut -c1-200
