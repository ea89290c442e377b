#!/bin/bash
# compute-sanitizer over a small-shape pass of every kernel class (scripts/sanitize_driver.py).  Run under gpurun:
#   gpurun --timeout 1500 -- scripts/sanitize.sh
# Logs land in gpurun_out/sanitize_<tool>.log; the summaries are copied to profiles/ by hand.
OUT=gpurun_out
python scripts/sanitize_driver.py > $OUT/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/sanitize_plain.log; exit 1; }
# our own detector first (always available): NaN-filled guard zones around every device buffer, see api_core.cu
GEORDD_GUARD=1 python scripts/sanitize_driver.py > $OUT/sanitize_guard.log 2>&1; echo "guard mode rc=$? : $(tail -1 $OUT/sanitize_guard.log)"
for tool in memcheck initcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_driver.py > $OUT/sanitize_$tool.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitize_$tool.log | tail -1)"
done
