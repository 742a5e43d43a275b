"""the N=8 bench workload, rank by rank, on one GPU: every rank's query batches through scan jobs + fetch (what bench.py --gpus 8 does per rank)"""
import faulthandler, os, sys, time
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from mmseqs2_b200 import Context, SubMatrix, synth
mat, pb = bench.load_matrix()
sm = SubMatrix(mat, pb)
bg = synth.background(pb)
res, off, _ = bench.make_scan_workload(0, int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 16, 2)
ctx = Context(0)
ctx.load_db(res, off, 21)
for rank in range(8):
    rng_q = np.random.default_rng(1000 + rank)
    qres, qoff = synth.random_seqs(rng_q, 32, bg, mean=350.0, sigma=35.0, lo=200, hi=500, normal=True)
    qs = synth.split(qres, qoff)
    lens = [len(q) for q in qs]
    batches = [[sm.ssw_query(q) for q in qs[b * 16:(b + 1) * 16]] for b in range(2)]
    jobs = [ctx.scan_job(bt, 15, 300) for bt in batches] + [ctx.scan_job(batches[0], 15, 300)]
    t0 = time.perf_counter()
    jobs[0].run(); jobs[1].run()
    for s in range(6):
        h, nh, _ = jobs[s % 3].fetch()
        if s + 2 < 6:
            jobs[(s + 2) % 3].run()
    ctx.sync()
    print("rank", rank, "lens", min(lens), max(lens), "ok %.3f s" % (time.perf_counter() - t0), "hits", int(nh.sum()), flush=True)
    for j in jobs:
        j.close()
ctx.close()
