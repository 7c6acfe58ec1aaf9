#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/group.log; : > $L
T=tools/fa_selftest
for g in 1 4 8 16 32 128; do
  echo "##### FA_B200_GROUP_HEADS=$g" >> $L
  for args in "4 32 8192 128 1 1 0 S 20" "1 32 16384 128 1 1 0 S 10" "16 16 2048 128 1 1 0 S 20" "2 16 4096 64 1 1 0 S 20"; do
    FA_B200_GROUP_HEADS=$g timeout 200 $T attn $args >> $L 2>&1; echo "exit=$?" >> $L
  done
done
grep -E "#####|FAIL|TIMING|exit=[1-9]" $L | cut -c1-200
