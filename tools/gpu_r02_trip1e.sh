#!/bin/bash
# round-2 trip 1e: full GPU suite, backward timing (slot-ring kernels), per-kernel times of the backward under ncu
TAG=${1:-r02e}
mkdir -p gpurun_out
L=gpurun_out/trip1e_$TAG.log; : > $L
timeout 900 python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit=$?" >> $L; tail -16 gpurun_out/pytest_gpu_$TAG.log >> $L
timeout 300 python tools/bwd_time.py >> $L 2>&1; echo "bwd_time exit=$?" >> $L
BW="python tools/bwd_time.py --one"
$BW > gpurun_out/plain_bwd.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fa_bwd -c 6 --csv --log-file gpurun_out/bwd_launches_$TAG.csv $BW > gpurun_out/ncu_bwd.log 2>&1
echo "ncu bwd exit=$?" >> $L
grep -E "fa_bwd" gpurun_out/bwd_launches_$TAG.csv | cut -d, -f5,13-15 | cut -c1-200 >> $L
cat $L | cut -c1-300 | tail -80
