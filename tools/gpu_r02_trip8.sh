#!/bin/bash
# full GPU suite on the in-tree library + the default bench line
TAG=${1:-r02i}
mkdir -p gpurun_out
L=gpurun_out/trip8_$TAG.log; : > $L
timeout 1200 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit=$?" >> $L; tail -15 gpurun_out/pytest_gpu_$TAG.log >> $L
timeout 600 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "bench exit=$?" >> $L
python - <<'PY' >> $L
import json,sys
d=json.loads(open('gpurun_out/bench_c3_%s.json' % sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/bench_c3_r02i.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')})
print('roofline', d['roofline']['frac'], d['roofline']['launch_ms_mean'], d['roofline']['launch_ms_min'])
print('e2e', d.get('e2e'))
for k in ('c2','c4','backward_c3'):
    print(k, d.get(k))
PY
cut -c1-400 $L
