#!/bin/bash
# CTA-0 timeline over its first four work items at c2 (a -DFA_TRACE build whose events are indexed by 8 * item + step)
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/build/trc
FA_B200_TRACE=gpurun_out/trace_c2_items_Z.txt timeout 100 tools/fa_selftest attn 8 16 1024 64 0 0 0 Z 0 > /dev/null 2>&1; echo "trace exit=$?"
python - <<'PY'
names = ["s0_ready", "s0_ld", "s0_pA", "s0_pB", "s1_ready", "s1_ld", "s1_pA", "s1_pB","pv0A", "pv0B", "qk0", "pv1A", "pv1B", "qk1", "s0_max", "s1_max"]
rows=[[int(x) for x in l.split()] for l in open('gpurun_out/trace_c2_items_Z.txt') if l.strip() and l[0].isdigit()]
t0=min(v for r in rows for v in r if v>0)
print("row(item*8+step) " + " ".join("%9s"%n for n in names[:8]))
for j,r in enumerate(rows[:32]):
    print("%2d (item %d step %d) "%(j,j//8,j%8) + " ".join("%9d"%((v-t0) if v else -1) for v in r[:8]))
PY
