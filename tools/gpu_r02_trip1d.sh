#!/bin/bash
# round-2 trip 1d: stream-memop probe (unbuffered), ring protocol emulation (event-based), c4 backward test, ncu of the new backward kernels
TAG=${1:-r02d}
mkdir -p gpurun_out
L=gpurun_out/trip1d_$TAG.log; : > $L
echo "##### memops probe" >> $L
timeout 25 tools/micro/memops_probe > gpurun_out/memops_probe_$TAG.log 2>&1; echo "probe exit=$?" >> $L; cat gpurun_out/memops_probe_$TAG.log >> $L
echo "##### ring emulation" >> $L
for cfg in "2 1" "2 0" "4 1" "3 1"; do
  timeout 90 python tests/_ring_emul.py $cfg 3 >> $L 2>&1; echo "emul $cfg exit=$?" >> $L
done
echo "##### ring + backward tests" >> $L
timeout 600 python -m pytest tests/test_backward_gpu.py tests/test_ring_gpu.py -q -m gpu > gpurun_out/pytest_bwd_ring_$TAG.log 2>&1; echo "pytest exit=$?" >> $L; tail -15 gpurun_out/pytest_bwd_ring_$TAG.log >> $L
BW="python tools/bwd_time.py --one"
$BW > gpurun_out/plain_bwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fa_bwd -s 2 -c 2 -o gpurun_out/prof_bwd_$TAG $BW > gpurun_out/ncu_bwd.log 2>&1
echo "ncu bwd exit=$?" >> $L
cat $L | cut -c1-300 | tail -120
