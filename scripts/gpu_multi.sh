#!/bin/bash
# scripts/gpu_multi.sh N -- multi-GPU evidence on an N-GPU box: the multi-device handle's tests, bench at 1..N, optionally the configs[3]/[4] script
N=${1:-2}
export PYTHONFAULTHANDLER=1 NCCL_DEBUG=WARN
mkdir -p gpurun_out
free -g | head -2; df -h /dev/shm | tail -1; cat /sys/fs/cgroup/memory.max 2>/dev/null
if [ "$4" != quick ]; then python -m pytest tests/test_multi_device.py -x -q -m gpu 2>&1 | tail -3; fi
for n in $([ "$4" = quick ] || echo 1) $N; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-secondary --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n value %.0f ms_per_step %.2f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    import re
    err = open("gpurun_out/bench_n$n.err").read()
    i = err.find("Fatal Python error")
    print("N=$n failed", e); print(err[i:i + 3000] if i >= 0 else err[-1500:])
PY
done
if [ "$2" = configs ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/multi_gpu_configs.py $3 2> gpurun_out/configs.err | tail -c 3000
  tail -5 gpurun_out/configs.err
fi
