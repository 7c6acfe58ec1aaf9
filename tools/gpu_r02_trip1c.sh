#!/bin/bash
# round-2 trip 1c: stream-memop probe, ring protocol emulation with diagnostics, new backward kernels (tests + timing, v1 vs v2)
TAG=${1:-r02c}
mkdir -p gpurun_out
L=gpurun_out/trip1c_$TAG.log; : > $L
echo "##### memops probe" >> $L
timeout 60 tools/micro/memops_probe >> $L 2>&1; echo "probe exit=$?" >> $L
echo "##### ring emulation world 2" >> $L
timeout 100 python tests/_ring_emul.py 2 0 2 >> $L 2>&1; echo "emul exit=$?" >> $L
timeout 100 python tests/_ring_emul.py 2 1 2 >> $L 2>&1; echo "emul exit=$?" >> $L
echo "##### backward tests" >> $L
timeout 600 python -m pytest tests/test_backward_gpu.py -q -m gpu > gpurun_out/pytest_bwd_$TAG.log 2>&1; echo "pytest bwd exit=$?" >> $L; tail -15 gpurun_out/pytest_bwd_$TAG.log >> $L
echo "##### backward timing, v2 kernels (in-tree lib)" >> $L
timeout 300 python tools/bwd_time.py >> $L 2>&1; echo "bwd_time exit=$?" >> $L
cat $L | cut -c1-300 | tail -120
