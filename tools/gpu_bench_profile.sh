#!/bin/bash
# bench (c3,c4,c2) + ncu launch list + ncu full capture; tag = $1
TAG=${1:-r01}
mkdir -p gpurun_out
L=gpurun_out/bench_$TAG.log
: > $L
for w in c3 c4 c2; do
  timeout 900 python bench.py --workload $w $( [ $w != c3 ] && echo --no-cpu-baseline ) > gpurun_out/bench_${w}_$TAG.json 2>> $L; echo "$w exit=$?" >> $L
done
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?" >> $L
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -o gpurun_out/prof_c3_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?" >> $L
cat $L
python - <<PY
import json
for w in ('c3','c4','c2'):
    d=json.load(open('gpurun_out/bench_%s_$TAG.json'%w))
    print(w, round(d['value'],1), round(d['ms_per_step'],4), d['clocks'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), round(d['roofline']['launch_ms_min'],4), 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
PY
