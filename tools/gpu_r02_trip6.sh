#!/bin/bash
# softmax-loop microbenchmark + CTA timelines of two library variants on zero and Set-S inputs
mkdir -p gpurun_out; L=gpurun_out/trip6.log; : > $L
timeout 120 tools/micro/softmax_loop >> $L 2>&1; echo "micro exit=$?" >> $L
T=tools/fa_selftest
for v in "$@"; do
  export LD_LIBRARY_PATH=$PWD/build/$v
  for set in Z S; do
    FA_B200_TRACE=gpurun_out/trace_${v}_$set.txt timeout 200 $T attn 4 32 8192 128 1 0 0 $set 0 > /dev/null 2>&1; echo "trace $v $set exit=$?" >> $L
    python tools/trace_report.py gpurun_out/trace_${v}_$set.txt 4 2>&1 | tail -16 >> $L
  done
done
cat $L | cut -c1-220
