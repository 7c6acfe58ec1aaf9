#!/bin/bash
# round-2 trip 4: fast softmax path (no row-max pass): parity suite on the default library, then interleaved A/B
TAG=${1:-r02h}
mkdir -p gpurun_out
L=gpurun_out/trip4_$TAG.log; : > $L
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x > gpurun_out/pytest_parity_$TAG.log 2>&1; echo "pytest exit=$?" >> $L; tail -25 gpurun_out/pytest_parity_$TAG.log >> $L
AB_EXTRA="4 32 8192 128 1 0 0 S 20" bash tools/gpu_ab.sh 2 fs0 fs1 fs1e1 >> $L 2>&1
cut -c1-250 $L | tail -90
