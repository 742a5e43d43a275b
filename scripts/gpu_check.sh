#!/bin/bash
# scripts/gpu_check.sh -- what one gpurun call does during development: parity tests, smoke, microbenchmark.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 120 ./profiles/microbench/int_pipe_rate > gpurun_out/int_pipe_rate.txt 2>&1
cat gpurun_out/int_pipe_rate.txt
