#!/bin/bash
# compute-sanitizer over the 20-residue parity case (memcheck, then racecheck on shared memory); logs under gpurun_out/
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/sanitize_$tool.log python tools/sanitize_case.py > gpurun_out/sanitize_$tool.out 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitize_$tool.log
done
