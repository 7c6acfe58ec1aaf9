#!/bin/bash
# A/B of the d<=64 "P in its own TMEM columns" variant: parity on d=64/32 shapes, then timings for both builds
mkdir -p gpurun_out; L=gpurun_out/sepP.log; : > $L
T=tools/fa_selftest
export LD_LIBRARY_PATH=$PWD/build/sepP1
echo "##### parity (sepP1)" >> $L
for a in "1 1 128 64 1 0" "1 2 777 64 0 1" "3 50 300 64 1 0" "1 16 1024 32 0 1" "2 3 1000 64 1 1 1300" "1 2 900 64 0 1 300" "1 1 8192 64 1 0" "1 2 4096 64 1 1" "2 4 2048 32 1 0" "1 2 1 64 1 1" "8 16 1024 64 0 0 0 R"; do
  timeout 120 $T attn $a >> $L 2>&1; echo "exit=$?" >> $L
done
for v in sepP0 sepP1 sepP0 sepP1; do
  echo "##### variant $v" >> $L
  export LD_LIBRARY_PATH=$PWD/build/$v
  for args in "8 16 1024 64 0 0 0 S 30" "2 16 4096 64 1 1 0 S 30" "4 16 8192 64 1 0 0 S 20" "4 32 4096 64 1 1 0 S 20" "1 16 1024 32 0 1 0 S 30" "4 16 8192 32 1 0 0 S 20"; do
    timeout 200 $T attn $args >> $L 2>&1; echo "exit=$?" >> $L
  done
done
grep -E "#####|FAIL|PASS|TIMING|exit=[1-9]" $L | cut -c1-230
