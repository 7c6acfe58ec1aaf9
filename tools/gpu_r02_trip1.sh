#!/bin/bash
# round-2 trip 1 (one B200): full GPU test-suite, smoke, bench (c3 + sub-records), comparators, backward timing,
# topology, then ncu: launch list of bench, full capture of the c2 (d=64) forward kernel and of both backward kernels.
TAG=${1:-r02a}
mkdir -p gpurun_out
L=gpurun_out/trip1_$TAG.log; : > $L
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $L 2>&1
(nvidia-smi topo -m; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; ls /sys/devices/system/node | tr '\n' ' ') > gpurun_out/topology_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit=$?" >> $L; tail -25 gpurun_out/pytest_gpu_$TAG.log >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "smoke exit=$?" >> $L
timeout 900 python bench.py > gpurun_out/bench_c3_$TAG.json 2>> $L; echo "bench exit=$?" >> $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> $L; echo "bench ref exit=$?" >> $L
timeout 900 python tools/comparators.py --out gpurun_out/comparators_$TAG > gpurun_out/comparators_$TAG.log 2>&1; echo "comparators exit=$?" >> $L
timeout 300 python tools/bwd_time.py >> $L 2>&1; echo "bwd_time exit=$?" >> $L
# ---- ncu (one tool per call; each command first exits 0 without ncu)
C2="python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-sub"
$C2 > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -o gpurun_out/prof_c2_$TAG $C2 > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 exit=$?" >> $L
BW="python tools/bwd_time.py --one"
$BW > gpurun_out/plain_bwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fa_bwd_sm100 -s 2 -c 2 -o gpurun_out/prof_bwd_$TAG $BW > gpurun_out/ncu_bwd.log 2>&1
echo "ncu bwd exit=$?" >> $L
C3="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-sub"
$C3 > gpurun_out/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $C3 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?" >> $L
cat $L | cut -c1-400 | tail -120
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c3_$TAG.json').read().strip().splitlines()[-1])
print('c3', round(d['value'],1), round(d['ms_per_step'],4), d['clocks'], 'e2e', d['e2e'] and (round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), round(d['e2e']['device_span_ms_per_step'],2)), 'roof', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], 'tmap', d.get('tmap_cache'))
for k in ('c2','c4'):
    if k in d: print(k, {x: (round(v,4) if isinstance(v,float) else v) for x,v in d[k].items() if x in ('launch_ms_mean','launch_ms_min','tflops','frac_of_measured_tensor_peak','hbm_gbs_algorithmic','frac_of_measured_hbm_peak')})
PY
cat gpurun_out/comparators_$TAG.md 2>/dev/null
