#!/bin/bash
# quick parity list + timing sweep for the library variants given as arguments (dirs under build/)
mkdir -p gpurun_out; L=gpurun_out/check.log; : > $L
T=tools/fa_selftest
for a in "1 1 128 128 1 0" "1 2 1000 128 1 1" "1 2 777 64 0 1" "1 2 900 128 0 1 300" "2 200 520 128 1 1" "3 50 300 64 1 0" "1 16 1024 32 0 1" "1 2 1024 128 1 0 0 R"; do
  timeout 120 $T attn $a >> $L 2>&1; echo "exit=$?" >> $L
done
grep -E "FAIL|exit=[1-9]" $L | cut -c1-200
bash tools/gpu_sweep2.sh "$@"
