#!/bin/bash
# Runs every test_conv case in its own process with a timeout (a deadlocked pipeline must not block the rest).
cd "$(dirname "$0")/../.."
BIN=tests/csrc/_bin/test_conv
N=$($BIN -1)
fail=0
for i in $(seq 0 $((N-1))); do
  timeout 60 $BIN $i
  rc=$?
  if [ $rc -ne 0 ]; then echo "[case $i] exit code $rc"; fail=1; fi
done
exit $fail
