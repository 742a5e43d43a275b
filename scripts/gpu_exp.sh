#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nucl.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python - <<'PY' > gpurun_out/nucl_perf.log 2>&1
import sys, time
sys.path.insert(0, '.')
import numpy as np
from mmseqs2_b200 import Context, synth
rng = np.random.default_rng(4)
nt = [synth.nucl_genome(rng, 30000) for _ in range(200)]
reads, tasks = synth.nucl_reads(rng, nt, 200000, 150, 0.02, 0.002)
td, to = synth.pack(nt)
ctx = Context(0); ctx.load_db(td, to, 5)
ctx.nucl_align(reads[:2000], tasks[:2000], decode=False); packed = synth.pack(reads)
for rep in range(3):
    t0 = time.perf_counter(); out, _, _ = ctx.nucl_align(packed, tasks, decode=False); dt = time.perf_counter() - t0
    print("e2e %.1f ms  %.0f aln/s   kernel %.1f ms" % (dt * 1e3, len(reads) / dt, ctx.last_kernel_ms))
PY
cat gpurun_out/nucl_perf.log
