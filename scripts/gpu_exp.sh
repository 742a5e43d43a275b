#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:"ungapped_scan_kernel" -s 4 -c 2 -o gpurun_out/prof_scan_full \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_scan_full.log 2>&1
tail -2 gpurun_out/ncu_scan_full.log | cut -c1-300
ls -la gpurun_out/prof_scan_full.ncu-rep
