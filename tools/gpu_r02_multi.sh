#!/bin/bash
# multi-GPU trip: ring tests with real processes, then the driver's own command line for bench.py --gpus N
N=${1:-2}; TAG=${2:-r02_n$N}
mkdir -p gpurun_out
L=gpurun_out/multi_$TAG.log; : > $L
nvidia-smi topo -m > gpurun_out/topology_$TAG.txt 2>&1
if [ "$3" != "notest" ]; then
timeout 600 python -m pytest tests/test_ring_gpu.py -q -m gpu -x > gpurun_out/pytest_ring_$TAG.log 2>&1; echo "pytest ring exit=$?" >> $L; tail -6 gpurun_out/pytest_ring_$TAG.log >> $L
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "bench exit=$?" >> $L
tail -5 gpurun_out/bench_c3_$TAG.err >> $L
cat $L | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c3_$TAG.json').read().strip().splitlines()[-1])
    print('headline', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', d['e2e'] and round(d['e2e']['value'],1))
    for k in ('strong_c3','ring_c5'):
        if k in d:
            r=dict(d[k]); r.pop('timeline_rank0_ms',None)
            print(k, json.dumps(r)[:1500])
    if 'ring_c5' in d and 'timeline_rank0_ms' in d['ring_c5']: print(json.dumps(d['ring_c5']['timeline_rank0_ms'])[:1500])
except Exception as e: print('parse failed', e)
PY
