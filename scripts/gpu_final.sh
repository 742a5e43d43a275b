#!/bin/bash
# scripts/gpu_final.sh -- evidence pass: DRAM traffic of the scan launches on the real config, the final bench line, the reference arm
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:ungapped_scan_kernel -c 5 -o gpurun_out/prof_scan_full \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_scan_full.log 2>&1
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cut -c1-300 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
cut -c1-400 gpurun_out/bench_reference_arm.json
ls -la gpurun_out/
