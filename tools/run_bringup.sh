#!/bin/bash
# bring-up driver for gpurun: all masks; on failure of the simplest case, a compute-sanitizer pass for the PC of the fault
mkdir -p gpurun_out
python tools/tile_bringup.py --batch ${2:-3} --masks "${1:-16,32,1,2,4,8,96,224,255}" > gpurun_out/bringup.log 2>&1
grep -E "^=====|RESULT|SUMMARY|BAD|RfError" gpurun_out/bringup.log | cut -c1-300
if ! grep -q "RESULT PASS" gpurun_out/bringup.log; then
  RF_TILE_MASK=16 timeout 300 compute-sanitizer --tool memcheck --print-limit 8 python tools/tile_bringup.py --child 16 --batch 1 > gpurun_out/sanitizer.log 2>&1
  grep -v "^step" gpurun_out/sanitizer.log | head -60
fi
