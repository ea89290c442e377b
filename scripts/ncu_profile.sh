#!/bin/bash
# Profiling recipe (B200_PROFILING.md): launch list of the bench command, then --set full captures of the dominant
# kernels.  Run under gpurun AFTER the same command exited 0 without ncu; results land in gpurun_out/ and the
# summaries are copied to profiles/ by hand (scripts/ncu_summarize.py).
set -u
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --draws 20000"
$CMD > $OUT/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/ncu_plain.log; exit 1; }
tail -c 600 $OUT/ncu_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
# the fused forward/backward solve sequence: second of the three step launches
ncu --set full --clock-control none --import-source on -k regex:solve_seq -s 1 -c 1 -o $OUT/prof_seq $CMD > $OUT/ncu_seq.log 2>&1
echo "seq capture rc=$?"
# triangular multiply = solve_dataflow_kernel<2, false>: second step launch (gpurun_out/ is limited to 64 MiB: the
# remaining captures go without source import)
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:solve_dataflow_kernel<\(int\)2' -s 1 -c 1 -o $OUT/prof_trmm $CMD > $OUT/ncu_trmm.log 2>&1
echo "trmm capture rc=$?"
ncu --set full --clock-control none -k regex:potrf_dataflow -c 3 -o $OUT/prof_potrf $CMD > $OUT/ncu_potrf.log 2>&1
echo "potrf capture rc=$?"
ncu --set full --clock-control none -k regex:solve_narrow -c 6 -o $OUT/prof_narrow $CMD > $OUT/ncu_narrow.log 2>&1
echo "narrow capture rc=$?"
ls -la $OUT
