#!/bin/bash
# round-end style validation on one B200: GPU tests, smoke, benches (c3,c4,c2), ncu launch list + full capture,
# the re-hosted reference cuda_fa1 driver next to the original, backward timing.  (compute-sanitizer is closed on this pool.)
TAG=${1:-final}
mkdir -p gpurun_out
L=gpurun_out/final_$TAG.log; : > $L
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit=$?" >> $L; tail -3 gpurun_out/pytest_gpu_$TAG.log >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "smoke exit=$?" >> $L
if [ "$2" != "nobench" ]; then bash tools/gpu_bench_profile.sh $TAG >> $L 2>&1; fi
echo "##### reference cuda_fa1 driver (main.cu) re-hosted on libfa_b200.so vs the original, same box" >> $L
for a in "1 8 512 64 4096 50" "8 16 1024 64 16384 50" "1 32 8192 128 16384 5"; do
  echo "--- main_b200 $a" >> $L; timeout 300 oracle/_ref/main_b200 $a 2>&1 | grep -v "^Error at" >> $L; echo "exit=$?" >> $L
done
for a in "1 8 512 64 4096 50" "8 16 1024 64 16384 20"; do
  echo "--- main_ref $a" >> $L; timeout 300 oracle/_ref/main_ref $a 2>&1 | grep -v "^Error at" >> $L; echo "exit=$?" >> $L
done
echo "##### backward timing (tools/bwd_time.py)" >> $L
timeout 200 python tools/bwd_time.py >> $L 2>&1; echo "exit=$?" >> $L
cat $L | cut -c1-250 | tail -150
