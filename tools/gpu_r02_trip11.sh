#!/bin/bash
# uniform-register hints (warp index / item ids through a lane-0 shuffle): full GPU suite + backward timing
mkdir -p gpurun_out; L=gpurun_out/trip11.log; : > $L
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_trip11.log 2>&1; echo "pytest exit=$?" >> $L; tail -4 gpurun_out/pytest_gpu_trip11.log >> $L
timeout 300 python tools/bwd_time.py --zeros >> $L 2>&1
timeout 300 python tools/bwd_time.py >> $L 2>&1
cat $L | cut -c1-200
