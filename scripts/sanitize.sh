#!/bin/bash
# compute-sanitizer passes over the tiny smoke scene (memcheck, racecheck, initcheck, synccheck).
set -u
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python __graft_entry__.py --smoke 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|error" | head -8
done
