#!/bin/bash
# Round-2 profiling recipe (B200_PROFILING.md): launch list of the bench command, then --set full captures of the dominant
# kernels, converted to CSV on the box (gpurun_out/ is limited to 64 MiB).  Run under gpurun AFTER the same command exited
# 0 without ncu.  Summaries are copied to profiles/ by scripts/ncu_summarize_r02.py.
set -u
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra"
$CMD > $OUT/r02_ncu_plain.json 2> $OUT/r02_ncu_plain.err || { echo "plain run failed"; tail -5 $OUT/r02_ncu_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/r02_launches.csv $CMD > $OUT/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
cap() {   # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -o $OUT/r02_$1 $CMD > $OUT/r02_ncu_$1.log 2>&1
  echo "$1 capture rc=$?"
  ncu -i $OUT/r02_$1.ncu-rep --page raw --csv > $OUT/r02_$1_raw.csv 2>/dev/null
  ncu -i $OUT/r02_$1.ncu-rep --page source --csv > $OUT/r02_$1_source.csv 2>/dev/null
  rm -f $OUT/r02_$1.ncu-rep
}
cap seq 'solve_seq_kernel' 3 1
cap trmm 'solve_dataflow_kernel<\(int\)2' 3 1
cap potrf 'potrf_dataflow_kernel' 0 3
cap cov 'cov2_kernel' 0 3
ls -la $OUT | grep r02_
