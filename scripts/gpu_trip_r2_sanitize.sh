#!/bin/bash
# compute-sanitizer memcheck over one tiny forward + backward of every kernel instantiation family (1 GPU).
mkdir -p gpurun_out
timeout 100 python benchmark/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain run exit $?"; tail -5 gpurun_out/sanitize_plain.log
timeout 420 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 7 python benchmark/sanitize_target.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "ok$|ERROR SUMMARY|Invalid|Error|error" gpurun_out/sanitize_memcheck.log | head -20
