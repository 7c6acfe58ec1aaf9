#!/bin/bash
# round-2 trip 3: tail-split tests, c2 / sharded-c3 timing with and without the split
TAG=${1:-r02g}
mkdir -p gpurun_out
L=gpurun_out/trip3_$TAG.log; : > $L
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "split or full_size_c2 or shard" > gpurun_out/pytest_split_$TAG.log 2>&1; echo "pytest exit=$?" >> $L; tail -15 gpurun_out/pytest_split_$TAG.log >> $L
T=tools/fa_selftest
for r in 1 2; do
for args in "8 16 1024 64 0 0 0 S 50" "1 16 8192 128 1 0 0 S 30" "4 32 8192 128 1 0 0 S 20"; do
  echo "## split on: $args" >> $L;  timeout 200 $T attn $args >> $L 2>&1; echo "exit=$?" >> $L
  echo "## split off: $args" >> $L; FA_NO_SPLIT=1 timeout 200 $T attn $args >> $L 2>&1; echo "exit=$?" >> $L
done
done
timeout 300 python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_$TAG.json 2>gpurun_out/bench_c2_$TAG.err; echo "bench c2 exit=$?" >> $L
grep -E "##|FAIL|TIMING|exit=|passed|failed" $L | cut -c1-300
