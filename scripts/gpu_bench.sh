#!/bin/bash
# scripts/gpu_bench.sh -- parity tests, smoke, bench line, ncu launch list, full ncu captures of the dominant kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json | cut -c1-600; tail -5 gpurun_out/bench.err
KREGEX='regex:ungapped_scan_kernel|sw16_kernel|sw32_kernel|topk_select_kernel|pad_profile_kernel|diag_score_kernel|nucl_align_kernel|sw_backtrace_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu --db-seqs 200000 --sw-queries 128 --sw-targets 128 --nucl-reads 20000 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sw16_kernel" -c 3 -o gpurun_out/prof_sw \
   python scripts/sw_perf.py > gpurun_out/ncu_sw.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ungapped_scan_kernel" -c 2 -o gpurun_out/prof_scan \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --db-seqs 200000 > gpurun_out/ncu_scan.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"nucl_align_kernel" -c 1 -o gpurun_out/prof_nucl \
   python bench.py --steps 1 --warmup 1 --no-cpu --db-seqs 20000 --sw-queries 8 --sw-targets 8 --nucl-reads 20000 > gpurun_out/ncu_nucl.log 2>&1
ls -la gpurun_out/
