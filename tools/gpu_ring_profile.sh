#!/bin/bash
# ring attention timeline at the per-step sizes of c5 on 8 GPUs, emulated with NG GPUs: N = 16384 * NG
NG=${1:-4}
mkdir -p gpurun_out
cat > /tmp/ring_prof.py <<'PY'
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
import flash_attention_impls_b200 as fa
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); torch.cuda.set_device(rank)
dev=torch.device("cuda",rank); dist.init_process_group("nccl", device_id=dev)
nl=16384
g=torch.Generator(device=dev); g.manual_seed(rank)
q=torch.randn((1,32,nl,128),generator=g,device=dev).bfloat16(); k=torch.randn_like(q); v=torch.rand_like(q)-0.5
for it in range(4):
    if it==3: os.environ["FA_B200_RING_PROFILE"]="1"
    dist.barrier(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); o,l=fa.ring_attention(q,k,v,causal=True); e1.record(); torch.cuda.synchronize()
    if rank==0: print("iter",it,"ms",e0.elapsed_time(e1), flush=True)
dist.destroy_process_group()
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29655 /tmp/ring_prof.py 2>&1 | grep -E "iter|ring profile rank [01]\]" | cut -c1-900 | tee gpurun_out/ring_profile_n$NG.log
