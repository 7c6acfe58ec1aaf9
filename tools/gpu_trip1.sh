#!/bin/bash
# Bring-up trip: descriptor unit tests, then attention parity on growing shapes. Every case runs in
# its own process under `timeout` so a trap/hang in one does not take the others down.
mkdir -p gpurun_out
L=gpurun_out/trip1.log
: > $L
run() { echo "### $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L 2>&1
T=tools/fa_selftest
run $T umma 128
run $T umma 64
run $T umma 128 1024 16384 2048
run $T umma 128 16384 1024 2048 0 1024
run $T attn 1 1 128 128 1 0
run $T attn 1 1 256 128 1 0
run $T attn 1 2 1024 128 1 0
run $T attn 1 2 1024 128 1 1
run $T attn 1 2 1024 128 0 0
run $T attn 1 1 128 64 0 0
run $T attn 1 2 1024 64 1 1
run $T attn 1 2 1000 128 1 0
run $T attn 1 2 1000 128 1 1
run $T attn 1 2 777 64 0 1
run $T attn 2 2 512 128 1 0 0 R
run $T attn 8 16 1024 64 0 0 0 S 20
run $T attn 4 32 8192 128 1 0 0 S 10
run $T attn 4 32 8192 128 1 1 0 S 10
run $T ref 1 2 512 64 4096 3
run $T ref 1 1 1024 128 16384 3
grep -E "RESULT|TIMING|exit=|watchdog|error" $L | tail -60
