#!/bin/bash
# round 2, call P: compute-sanitizer memcheck on the small end-to-end case (one sanitizer tool per call)
set -x
mkdir -p gpurun_out
timeout 300 python profiles/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/sanitize_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_case.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/sanitize_memcheck.log
grep -c "Invalid\|invalid\|Error" gpurun_out/sanitize_memcheck.log
