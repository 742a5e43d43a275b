#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python -c "
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print(d['value'], d['e2e']['value'])
print(d['secondary'].get('search_pipeline'))
print(d['secondary']['nucl_align'].get('kernel_ms'), d['secondary']['nucl_align'].get('e2e'))
"
