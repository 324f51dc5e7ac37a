#!/bin/bash
# compute-sanitizer over every kernel family of the library (tools/sanitize_driver.py); logs land in $1 (default gpurun_out/).
# Run on the GPU box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
rc=0
timeout 700 $CS --tool memcheck --leak-check no --error-exitcode 1 --log-file "$OUT/sanitizer_memcheck.log" python tools/sanitize_driver.py > "$OUT/sanitizer_memcheck.out" 2>&1 || rc=1
timeout 500 $CS --tool racecheck --racecheck-report all --error-exitcode 1 --log-file "$OUT/sanitizer_racecheck.log" python tools/sanitize_driver.py quick > "$OUT/sanitizer_racecheck.out" 2>&1 || rc=1
timeout 300 $CS --tool synccheck --error-exitcode 1 --log-file "$OUT/sanitizer_synccheck.log" python tools/sanitize_driver.py quick > "$OUT/sanitizer_synccheck.out" 2>&1 || rc=1
tail -n 3 "$OUT"/sanitizer_*.log
echo "sanitize rc=$rc"
exit $rc
