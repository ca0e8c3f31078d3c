#!/bin/bash
# Runs every test_conv case in its own process with a timeout (a deadlocked pipeline must not block the rest;
# two timeouts in a row stop the run: GPU time is budgeted).
cd "$(dirname "$0")/../.."
BIN=tests/csrc/_bin/test_conv
N=$($BIN -1)
fail=0
hangs=0
for i in $(seq 0 $((N-1))); do
  timeout ${CONV_CASE_TIMEOUT:-30} $BIN $i
  rc=$?
  if [ $rc -ne 0 ]; then echo "[case $i] exit code $rc"; fail=1; fi
  if [ $rc -eq 124 ]; then hangs=$((hangs+1)); else hangs=0; fi
  if [ $hangs -ge 2 ]; then echo "two cases in a row timed out: stopping"; break; fi
done
exit $fail
