#!/bin/bash
# pytest -m gpu, smoke, bench (both arms), then ncu launch list + one full capture of the top kernel.
mkdir -p gpurun_out
L=gpurun_out/trip2.log
: > $L
echo "### pytest -m gpu" >> $L
timeout 1200 python -m pytest tests -x -q -m gpu >> $L 2>&1; echo "exit=$?" >> $L
echo "### smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "exit=$?" >> $L
echo "### bench reference arm" >> $L
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 >> $L 2>&1; echo "exit=$?" >> $L
echo "### bench c3" >> $L
timeout 900 python bench.py > gpurun_out/bench_c3.json 2>> $L; echo "exit=$?" >> $L; cat gpurun_out/bench_c3.json >> $L
for w in c4 c2; do
  echo "### bench $w" >> $L
  timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2>> $L; echo "exit=$?" >> $L; cat gpurun_out/bench_$w.json >> $L
done
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
echo "### ncu" >> $L
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?" >> $L
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 2 -o gpurun_out/prof_c3 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?" >> $L
tail -5 gpurun_out/ncu_full.log >> $L
grep -E "passed|failed|error|exit=|smoke|impl|metric" $L | cut -c1-600 | tail -40
