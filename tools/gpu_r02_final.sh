#!/bin/bash
# round-2 validation on one B200: GPU tests, smoke, default bench line (with sub-records), reference arm, ncu launch
# list + full capture of the forward kernel, backward timing
TAG=${1:-r02final}
mkdir -p gpurun_out
L=gpurun_out/final_$TAG.log; : > $L
timeout 1200 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit=$?" >> $L; tail -12 gpurun_out/pytest_gpu_$TAG.log >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "smoke exit=$?" >> $L
timeout 900 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "bench exit=$?" >> $L
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_c3_$TAG.err; echo "bench reference exit=$?" >> $L
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-sub"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?" >> $L
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -o gpurun_out/prof_c3_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?" >> $L
python - $TAG <<'PY' >> $L
import json,sys
tag=sys.argv[1]
d=json.loads(open('gpurun_out/bench_c3_%s.json'%tag).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')})
print('roofline', d['roofline']['frac'], d['roofline']['launch_ms_mean'], d['roofline']['launch_ms_min'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for k in ('c2','c4','backward_c3','c3_zero_inputs'):
    print(k, {a:b for a,b in d.get(k,{}).items() if a in ('tflops','launch_ms_mean','launch_ms_min','frac_of_measured_tensor_peak','tflops_5_product_convention','ms_per_step','random_over_zero','hbm_gbs_algorithmic')})
print('cpu_baseline', d.get('cpu_baseline'))
PY
cut -c1-300 $L | tail -60
