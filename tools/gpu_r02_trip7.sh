#!/bin/bash
# softmax-tile sandbox, fast-path v2 variants on zero / Set-S inputs, timeline, parity of the in-tree library
mkdir -p gpurun_out; L=gpurun_out/trip7.log; : > $L
timeout 120 tools/micro/softmax_tile >> $L 2>&1; echo "sandbox exit=$?" >> $L
T=tools/fa_selftest
for r in 1 2; do
for v in fs0 fs2 fs0p3 fs2p3; do
  export LD_LIBRARY_PATH=$PWD/build/$v
  for args in "4 32 8192 128 1 0 0 Z 10" "4 32 8192 128 1 0 0 S 20" "4 32 8192 128 1 1 0 S 20" "4 32 8192 64 1 0 0 Z 10" "8 16 1024 64 0 0 0 S 30"; do
    echo "##### $v: $args" >> $L
    timeout 200 $T attn $args 2>&1 | grep -E "TIMING|FAIL|watchdog|error" >> $L
  done
done
done
export LD_LIBRARY_PATH=$PWD/build/tr2
for set in Z S; do
  FA_B200_TRACE=gpurun_out/trace_tr2_$set.txt timeout 200 $T attn 4 32 8192 128 1 0 0 $set 0 > /dev/null 2>&1; echo "trace tr2 $set exit=$?" >> $L
  python tools/trace_report.py gpurun_out/trace_tr2_$set.txt 4 2>&1 | tail -14 >> $L
done
unset LD_LIBRARY_PATH
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x > gpurun_out/pytest_parity_trip7.log 2>&1; echo "pytest exit=$?" >> $L; tail -8 gpurun_out/pytest_parity_trip7.log >> $L
cut -c1-220 $L
