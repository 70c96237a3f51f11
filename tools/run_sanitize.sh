#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_target.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|fp16 streams|int8|npp|exchange|Error|error" gpurun_out/sanitize_memcheck.log | head -30
