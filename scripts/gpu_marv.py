"""GPU box: the reference's Marv::scan (baseline/_ref) vs libb200align on the same 1M-sequence DB and queries."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from baseline import marv
from mmseqs2_b200 import Context, SubMatrix

db_seqs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 32
res, off, queries = bench.make_scan_workload(0, db_seqs, nq, 1)
mat, pb = bench.load_matrix()
sm = SubMatrix(mat, pb)
profs = [sm.ssw_query(q) for q in queries]
out = {"db_seqs": db_seqs, "queries": nq, "db_residues": int(off[-1])}
ctx = Context(0)
ctx.load_db(res, off, 21)
ctx.ungapped_scan(profs[:16], 15, 300)
t0 = time.perf_counter()
ours = []
for b in range(0, nq, 16):
    h, nh, _ = ctx.ungapped_scan(profs[b:b + 16], 15, 300)
    ours += [(h[i]["id"][:int(nh[i])].copy(), h[i]["score"][:int(nh[i])].copy()) for i in range(len(h))]
dt = time.perf_counter() - t0
cells = sum(len(q) for q in queries) * float(off[-1])
out["b200_e2e_gcups"] = cells / 1e9 / dt
ctx.close()
tb, hits = marv.time_tables(res, off, queries, [p.profile for p in profs])
out["libmarv"] = tb
# how the reference's GPU scores relate to the CPU-exact ones (informational: Marv does not saturate at 255-bias)
agree = {}
for name, lists in hits.items():
    same_top, n = 0, 0
    for (mi, ms), (oi, os_) in zip(lists, ours):
        keep = ms > 15
        a = set(zip(mi[keep].tolist(), ms[keep].tolist()))
        b = set(zip(oi.tolist(), os_.tolist()))
        same_top += len(a & b); n += max(len(a), len(b))
    agree[name] = same_top / max(1, n)
out["hit_overlap_with_b200"] = agree
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "marv.json"), "w"), indent=1)
