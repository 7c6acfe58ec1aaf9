#!/bin/bash
# round-2 trip 1b: full GPU suite (no -x, to see every failure), A/B of the P-publication split, backward timing
TAG=${1:-r02b}
mkdir -p gpurun_out
L=gpurun_out/trip1b_$TAG.log; : > $L
timeout 1500 python -m pytest tests -q -m gpu --durations=10 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit=$?" >> $L; tail -40 gpurun_out/pytest_gpu_$TAG.log >> $L
bash tools/gpu_ab.sh 2 base p3 > gpurun_out/ab_$TAG.txt 2>&1; cat gpurun_out/ab_$TAG.txt >> $L
timeout 300 python tools/bwd_time.py >> $L 2>&1; echo "bwd_time exit=$?" >> $L
cat $L | cut -c1-300 | tail -150
