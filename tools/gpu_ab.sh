#!/bin/bash
# interleaved A/B timing of library variants (dirs under build/): usage gpu_ab.sh ROUNDS v1 v2 ...
R=$1; shift
mkdir -p gpurun_out; L=gpurun_out/ab.log; : > $L
T=tools/fa_selftest
for r in $(seq 1 $R); do
  for v in "$@"; do
    echo "##### variant $v round $r" >> $L
    export LD_LIBRARY_PATH=$PWD/build/$v
    for args in "4 32 8192 128 1 1 0 S 20" "1 32 16384 128 1 1 0 S 10" "8 16 1024 64 0 0 0 S 30" ${AB_EXTRA:+"$AB_EXTRA"}; do
      timeout 200 $T attn $args >> $L 2>&1; echo "exit=$?" >> $L
    done
  done
done
grep -E "#####|FAIL|TIMING|exit=[1-9]" $L | cut -c1-200
