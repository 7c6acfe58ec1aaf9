#!/bin/bash
# all GPUs of the box loaded at once: the same fa_selftest timing per GPU, variants in turn (is a variant's rate different
# when the whole chassis draws power?)  usage: gpu_ab8.sh NGPU v1 v2 ...
N=$1; shift
mkdir -p gpurun_out; L=gpurun_out/ab8.log; : > $L
T=tools/fa_selftest
for r in 1 2; do
for v in "$@"; do
  for g in $(seq 0 $((N-1))); do
    ( export CUDA_VISIBLE_DEVICES=$g LD_LIBRARY_PATH=$PWD/build/$v; timeout 300 $T attn 4 32 8192 128 1 0 0 S 40 2>&1 | grep TIMING | sed "s/^/$v gpu$g: /" >> $L ) &
  done
  wait
done
done
sed 's/TIMING attn B=4 H=32 N=8192 d=128 bf16 causal=0: //' $L | sort | cut -c1-150
