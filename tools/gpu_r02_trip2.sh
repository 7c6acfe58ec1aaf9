#!/bin/bash
# round-2 trip 2: full GPU suite (incl. general head dims, backward strides / N_kv), backward timing
TAG=${1:-r02f}
mkdir -p gpurun_out
L=gpurun_out/trip2_$TAG.log; : > $L
timeout 900 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit=$?" >> $L; tail -40 gpurun_out/pytest_gpu_$TAG.log >> $L
timeout 300 python tools/bwd_time.py >> $L 2>&1; echo "bwd_time exit=$?" >> $L
cat $L | cut -c1-400 | tail -100
