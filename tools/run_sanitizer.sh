#!/bin/bash
# compute-sanitizer memcheck / racecheck / synccheck over tools/sanitize_k1.py (small shapes).
# Usage (GPU box): bash tools/run_sanitizer.sh [tool ...]   -> gpurun_out/sanitizer_<tool>.log
mkdir -p gpurun_out
tools=("$@")
[ ${#tools[@]} -eq 0 ] && tools=(memcheck racecheck synccheck)
rc=0
for t in "${tools[@]}"; do
  log=gpurun_out/sanitizer_$t.log
  echo "=== compute-sanitizer --tool $t" | tee "$log"
  timeout 900 compute-sanitizer --tool "$t" --error-exitcode 7 --print-limit 20 \
      python tools/sanitize_k1.py >> "$log" 2>&1
  code=$?
  echo "exit code $code" | tee -a "$log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_DRIVER_OK|exit code" "$log" | tail -5
  [ $code -ne 0 ] && rc=1
done
exit $rc
