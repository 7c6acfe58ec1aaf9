#!/bin/bash
NG=${1:-4}
mkdir -p gpurun_out
L=gpurun_out/ring_n$NG.log
: > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611"
timeout 900 python -m pytest tests/test_ring_gpu.py -x -q -m gpu >> $L 2>&1; echo "pytest exit=$?" >> $L
timeout 1200 $TR bench.py --gpus $NG --steps 5 --warmup 3 --workload c5 > gpurun_out/bench_c5_n$NG.json 2>> $L; echo "exit=$?" >> $L
cat gpurun_out/bench_c5_n$NG.json >> $L
grep -E "passed|failed|rror|exit=|metric" $L | cut -c1-330 | tail
