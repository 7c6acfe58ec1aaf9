#!/bin/bash
# backward: zero-input vs random timing, ncu --set full of the final kernels; forward c2: ncu --set full of the shared-S kernel
mkdir -p gpurun_out; L=gpurun_out/trip10.log; : > $L
timeout 300 python tools/bwd_time.py --zeros >> $L 2>&1; echo "exit=$?" >> $L
timeout 300 python tools/bwd_time.py >> $L 2>&1; echo "exit=$?" >> $L
BW="python tools/bwd_time.py --one"
$BW > gpurun_out/plain_bwd.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fa_bwd -s 2 -c 2 -o gpurun_out/prof_bwd_r02final $BW > gpurun_out/ncu_bwd.log 2>&1
echo "ncu bwd exit=$?" >> $L
C2="python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-sub"
$C2 > gpurun_out/plain_c2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -o gpurun_out/prof_c2_r02final $C2 > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 exit=$?" >> $L
cat $L | cut -c1-200
