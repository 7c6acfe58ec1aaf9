#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/v6.log
: > $L
T=tools/fa_selftest
run() { echo "### $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run $T attn 1 1 128 128 1 0
run $T attn 1 1 256 128 1 0
run $T attn 1 2 1024 128 1 0
run $T attn 1 2 1024 128 1 1
run $T attn 1 1 128 64 0 0
run $T attn 1 2 1000 128 1 1
run $T attn 1 2 777 64 0 1
run $T attn 1 2 900 128 0 1 300
run $T attn 3 50 300 64 1 0
run $T attn 2 200 520 128 1 1
run $T attn 4 32 8192 128 1 0 0 S 20
run $T attn 4 32 8192 128 1 1 0 S 20
run $T attn 8 16 1024 64 0 0 0 S 20
run $T attn 2 16 4096 64 1 1 0 S 20
grep -E "RESULT|TIMING|exit=[1-9]|watchdog|rror" $L | cut -c1-220
