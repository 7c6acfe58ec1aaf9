#!/bin/bash
# parity list for library variants (build/$v): usage gpu_variant_parity.sh v1 ...
mkdir -p gpurun_out
T=tools/fa_selftest
for v in "$@"; do
  L=gpurun_out/variant_$v.log; : > $L
  export LD_LIBRARY_PATH=$PWD/build/$v
  for a in "1 1 128 128 1 0" "1 2 1000 128 1 1" "1 2 777 64 0 1" "1 2 900 128 0 1 300" "2 200 520 128 1 1" "3 50 300 64 1 0" "1 16 1024 32 0 1" "1 1 8192 64 1 0" "2 3 1000 64 1 1 1300" "1 4 4096 128 1 0"; do
    timeout 60 $T attn $a >> $L 2>&1; echo "exit=$?" >> $L
  done
  echo "== $v: $(grep -c 'RESULT attn PASS' $L) pass, $(grep -c 'exit=[1-9]' $L) bad exits"
  grep -E "FAIL|exit=[1-9]|watchdog|error" $L | cut -c1-200
done
