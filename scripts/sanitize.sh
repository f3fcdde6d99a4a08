#!/bin/bash
# compute-sanitizer over scripts/sanitize_driver.py (every kernel, every polygon tier).  Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
# Writes gpurun_out/sanitize_<tool>.log.  NOTE (round 1): compute-sanitizer is closed on this GPU pool (the wrapper
# refuses to start it), so only the plain run of the driver -- every kernel and tier against the oracle -- is evidence.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
python scripts/sanitize_driver.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "driver fails without sanitizer"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
tail -2 gpurun_out/sanitize_plain.log
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
    extra=""
    [ "$tool" = racecheck ] && extra="--racecheck-report all"
    [ "$tool" = initcheck ] && extra="--track-unused-memory no"
    start=$(date +%s)
    timeout ${TOOL_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 20 \
        python scripts/sanitize_driver.py > gpurun_out/sanitize_$tool.log 2>&1
    rc=$?
    echo "== $tool rc=$rc $(( $(date +%s) - start ))s"
    grep -E "RESULT|ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -8
done
