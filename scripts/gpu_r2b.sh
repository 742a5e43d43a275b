#!/bin/bash
# round-2 evidence pass B: full GPU tests, nucleotide occupancy variants, align-step phase trace, ncu captures of the scan
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python scripts/nucl_perf.py 2>&1 | tail -4
SWQ=1024 SWT=1024 B200_TRACE=1 python scripts/align_step_perf.py > gpurun_out/align_trace.log 2>&1; grep -v "^\[b200 trace\] backtrace:" gpurun_out/align_trace.log | tail -30 | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-libmarv --sw-queries 128 --sw-targets 128 --nucl-reads 20000 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ungapped_scan_kernel -s 6 -c 5 -o gpurun_out/r02_prof_scan python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary > gpurun_out/ncu_scan.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
