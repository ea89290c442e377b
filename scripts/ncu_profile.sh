#!/bin/bash
# Round-1 profiling recipe (B200_PROFILING.md): launch list of the bench command, then one --set full capture of the
# dominant kernels.  Run under gpurun; results land in gpurun_out/ and summaries are copied to profiles/ by hand.
set -u
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --draws 20000"
$CMD > $OUT/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/ncu_plain.log; exit 1; }
tail -c 600 $OUT/ncu_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
# 14 matching launches belong to the setup (3 fits x (potrf, fwd, bwd), cliff face, Sigma factor, denominators)
ncu --set full --clock-control none --import-source on -k regex:solve_dataflow -s 14 -c 7 -o $OUT/prof_solve $CMD > $OUT/ncu_solve.log 2>&1
echo "solve capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:potrf_dataflow -s 2 -c 1 -o $OUT/prof_potrf $CMD > $OUT/ncu_potrf.log 2>&1
echo "potrf capture rc=$?"
ls -la $OUT
