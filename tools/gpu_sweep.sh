#!/bin/bash
# Perf sweep over library variants in build/lib_*: parity spot checks + timing of c3/c4/c2.
mkdir -p gpurun_out
L=gpurun_out/sweep.log
: > $L
T=tools/fa_selftest
for v in "$@"; do
  echo "##### variant $v" >> $L
  export LD_LIBRARY_PATH=$PWD/build/$v
  for args in "1 2 1000 128 1 1" "1 2 777 64 0 1" "1 2 900 128 0 1 300" "4 32 8192 128 1 0 0 S 20" "4 32 8192 128 1 1 0 S 20" "8 16 1024 64 0 0 0 S 20" "2 16 4096 64 1 1 0 S 20"; do
    timeout 120 $T attn $args >> $L 2>&1; echo "exit=$?" >> $L
  done
done
grep -E "#####|RESULT|TIMING|exit=[1-9]" $L | cut -c1-230
