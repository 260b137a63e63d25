#!/bin/bash
# Run every GPU test id in its own process (a device-side trap must not poison the rest); log per test.
mkdir -p gpurun_out
LOG=gpurun_out/gpu_tests.log
: > $LOG
ids=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::")
pass=0; fail=0
for id in $ids; do
  echo "=== $id" >> $LOG
  if timeout 300 python -m pytest "$id" -x -q -s 2>&1 | tail -25 >> $LOG; then :; fi
  if tail -3 $LOG | grep -q "passed"; then pass=$((pass+1)); echo "PASS $id"; else fail=$((fail+1)); echo "FAIL $id"; fi
done
echo "passed=$pass failed=$fail" | tee -a $LOG
