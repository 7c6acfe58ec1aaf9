#!/bin/bash
# one torchrun for parity+timing+timeline of both ring transports, then the official c5 bench line (peer transport)
NG=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
L=gpurun_out/ring_eval_n$NG.log
timeout 300 $TR --master-port 29611 tools/ring_eval.py > $L 2>&1; echo "ring_eval exit=$?" >> $L
timeout 240 $TR --master-port 29613 bench.py --gpus $NG --steps 5 --warmup 3 --workload c5 > gpurun_out/bench_c5_n${NG}_peer.json 2> gpurun_out/bench_c5_n${NG}_peer.err; echo "bench exit=$?" >> $L
grep -E "PARITY|TIMING|ring profile rank 0|exit=|Error|error" $L | cut -c1-700 | tail -30
cat gpurun_out/bench_c5_n${NG}_peer.json | cut -c1-600
