#!/bin/bash
# shared-S TMEM protocol: parity list on the in-tree library, then zero / Set-S timing of the variants given
mkdir -p gpurun_out; L=gpurun_out/trip9.log; : > $L
T=tools/fa_selftest
for a in "1 1 128 128 1 0" "1 2 1000 128 1 1" "1 2 777 64 0 1" "1 2 900 128 0 1 300" "2 200 520 128 1 1" "3 50 300 64 1 0" "1 16 1024 32 0 1" "1 2 1024 128 1 0 0 R" "1 1 8192 64 1 0" "2 3 1000 64 1 1 1300" "1 4 4096 128 1 0" "1 4 4096 128 0 1"; do
  timeout 60 $T attn $a >> $L 2>&1; echo "exit=$?" >> $L
  FA_STATS=1 timeout 60 $T attn $a >> $L 2>&1; echo "exit=$?" >> $L
done
grep -E "RESULT|FAIL|exit=[1-9]|watchdog|error" $L | cut -c1-220
if grep -q "exit=[1-9]" $L; then echo "PARITY FAILED - skipping timing"; exit 1; fi
for r in 1 2; do
for v in "$@"; do
  export LD_LIBRARY_PATH=$PWD/build/$v
  for args in "4 32 8192 128 1 0 0 Z 10" "4 32 8192 128 1 0 0 S 20" "4 32 8192 128 1 1 0 S 20" "4 32 8192 64 1 0 0 Z 10" "4 32 8192 64 1 0 0 S 20" "8 16 1024 64 0 0 0 S 30"; do
    echo "##### $v: $args" >> $L
    timeout 200 $T attn $args 2>&1 | grep -E "TIMING|FAIL|watchdog|error" | tee -a $L | cut -c1-200 | sed "s/^/$v: /"
  done
done
done
unset LD_LIBRARY_PATH
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x > gpurun_out/pytest_parity_trip9.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/pytest_parity_trip9.log
