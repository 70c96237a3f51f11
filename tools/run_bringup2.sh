#!/bin/bash
# bring-up + NPP dump in one GPU call
mkdir -p gpurun_out
python tools/npp_dump.py > gpurun_out/npp_dump.log 2>&1; tail -12 gpurun_out/npp_dump.log
python tools/tile_bringup.py --batch ${2:-8} --masks "${1:-255,96,16,15}" > gpurun_out/bringup.log 2>&1
grep -E "^=====|RESULT|SUMMARY|BAD|RfError" gpurun_out/bringup.log | cut -c1-200
