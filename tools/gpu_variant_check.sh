#!/bin/bash
# parity list for one library variant (build/$1), then interleaved A/B timing against build/base
V=$1
mkdir -p gpurun_out; L=gpurun_out/variant_$V.log; : > $L
T=tools/fa_selftest
export LD_LIBRARY_PATH=$PWD/build/$V
for a in "1 1 128 128 1 0" "1 2 1000 128 1 1" "1 2 777 64 0 1" "1 2 900 128 0 1 300" "2 200 520 128 1 1" "3 50 300 64 1 0" "1 16 1024 32 0 1" "1 2 1024 128 1 0 0 R" "1 1 8192 64 1 0" "2 3 1000 64 1 1 1300"; do
  timeout 60 $T attn $a >> $L 2>&1; echo "exit=$?" >> $L
done
grep -E "RESULT|FAIL|exit=[1-9]|watchdog|error" $L | cut -c1-200
if grep -q "exit=[1-9]" $L; then echo "PARITY FAILED - skipping timing"; exit 1; fi
bash tools/gpu_ab.sh 2 base $V
