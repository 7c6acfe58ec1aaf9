#!/bin/bash
# ncu --set full of the final forward kernel at c2 and c4
mkdir -p gpurun_out; L=gpurun_out/trip12.log; : > $L
for w in c2 c4; do
  C="python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-sub"
  $C > gpurun_out/plain_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -o gpurun_out/prof_${w}_r02n $C > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w exit=$?" >> $L
done
cat $L
