#!/bin/bash
# Multi-GPU trip: NCCL ring tests, (b,h)-sharded bench, ring-attention bench (c5).
NG=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/trip3_n$NG.log
: > $L
nvidia-smi -L >> $L 2>&1
echo "### pytest ring" >> $L
timeout 900 python -m pytest tests/test_ring_gpu.py -x -q -m gpu >> $L 2>&1; echo "exit=$?" >> $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611"
echo "### bench c3 x$NG" >> $L
timeout 900 $TR bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/bench_c3_n$NG.json 2>> $L; echo "exit=$?" >> $L; cat gpurun_out/bench_c3_n$NG.json >> $L
echo "### bench c4 x$NG" >> $L
timeout 900 $TR bench.py --gpus $NG --steps 20 --warmup 5 --workload c4 > gpurun_out/bench_c4_n$NG.json 2>> $L; echo "exit=$?" >> $L; cat gpurun_out/bench_c4_n$NG.json >> $L
echo "### bench c5 ring x$NG" >> $L
timeout 1200 $TR bench.py --gpus $NG --steps 5 --warmup 3 --workload c5 > gpurun_out/bench_c5_n$NG.json 2>> $L; echo "exit=$?" >> $L; cat gpurun_out/bench_c5_n$NG.json >> $L
grep -E "passed|failed|rror|exit=|metric" $L | cut -c1-700 | tail -30
