#!/bin/bash
# round-2 final single-GPU evidence: tests, smoke, bench (both arms), search wall-clock through the patched host
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_final.json").read().strip().splitlines()[-1])
print("value %.0f e2e %.0f frac %.3f smem_frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["binding_resource"]["frac"]))
s = d["secondary"]
print("sw16 %.0f frac %.3f | align_step %.0f aln/s identical %s | nucl %.0f | libmarv %s | cpu %s" % (
    s["sw_rescoring"]["value"], s["sw_rescoring"]["roofline"]["frac"], s["sw_rescoring"]["align_step"]["alignments_per_s"],
    s["sw_rescoring"]["align_step"].get("cpu_baseline", {}).get("records_identical_to_gpu"), s["nucl_align"]["value"],
    {k: round(v["gcups_wall"]) for k, v in s["libmarv"]["GCUPS"].items()}, {k: d["cpu_baseline"].get(k) for k in ("value", "cores", "scan_identical_to_gpu", "hit_lists_identical_to_gpu")}))
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference_arm.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_r2_reference_arm.json
B200_TRACE=1 timeout 900 python integration/search_wallclock.py --db-seqs 1000000 --queries 768 --cpu-queries 64 --work /tmp/sw --out gpurun_out/search_wallclock_1M_b.json 2>&1 | tail -c 2200
grep -h "libb200align" /tmp/sw/search_b200.log | tail -4
