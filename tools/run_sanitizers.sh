#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over tools/sanitize_smoke.py; summaries -> gpurun_out/sanitizer/
# usage: bash tools/run_sanitizers.sh [targets...]   (default: movegen trunk engine wide wide_engine)
out=gpurun_out/sanitizer
mkdir -p $out
targets=${@:-movegen trunk engine wide wide_engine}
for tool in memcheck racecheck synccheck; do
  for t in $targets; do
    log=$out/${tool}_${t}.log
    start=$(date +%s)
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py $t > $log 2>&1
    rc=$?
    echo "exit $rc after $(( $(date +%s) - start )) s" >> $log
    echo "== $tool $t: rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
  done
done
