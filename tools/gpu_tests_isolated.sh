#!/bin/bash
# Run every GPU test id in its own process with a timeout (a trapped kernel poisons its CUDA
# context), append results to gpurun_out/isolated.log.  Usage: tools/gpu_tests_isolated.sh [pytest -k expr]
mkdir -p gpurun_out
LOG=gpurun_out/isolated.log
: > $LOG
ids=$(python -m pytest tests -m gpu --collect-only -q ${1:+-k "$1"} 2>/dev/null | grep "::")
for id in $ids; do
  start=$(date +%s)
  timeout 240 python -m pytest "$id" -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/_one.log 2>&1
  rc=$?
  echo "[$rc] $(( $(date +%s) - start ))s $id" | tee -a $LOG
  if [ $rc -ne 0 ]; then grep -E "Error|error|assert|rel err|timed out|FAILED|failed" gpurun_out/_one.log | head -12 | sed 's/^/      /' | tee -a $LOG; fi
done
echo "passed: $(grep -c '^\[0\]' $LOG)  failed: $(grep -vc '^\[0\]' $LOG | head -1)" | tee -a $LOG
