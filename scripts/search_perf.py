"""scripts/search_perf.py -- phase times of one search batch (scan -> hit lists -> b200_align_batch); run with B200_TRACE=1."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mmseqs2_b200 import Context, SubMatrix  # noqa: E402
from mmseqs2_b200 import alignment as al  # noqa: E402


def main():
    mat, pb = bench.load_matrix()
    sm = SubMatrix(mat, pb)
    res, off, queries = bench.make_scan_workload(0, int(os.environ.get("DBSEQS", "300000")), 16, 2)
    ctx = Context(0)
    ctx.load_db(res, off, 21)
    evp = al.EvalueParams.defaults("blosum62.out", 11, 1, int(off[-1]))
    par = al.AlignParams(sw_mode=al.SCORE_COV_SEQID, eval_thr=1e-3)
    for rep in range(3):
        qs = queries[(rep % 2) * 16:(rep % 2) * 16 + 16]
        profs = [sm.ssw_query(q) for q in qs]
        t0 = time.perf_counter()
        h, nh, _ = ctx.ungapped_scan(profs, 15, 300)
        t1 = time.perf_counter()
        lists = [h[qi]["id"][:int(nh[qi])] for qi in range(16)]
        r, pool, n_aln = al.align_batch(ctx, sm, qs, lists, par, evp)
        t2 = time.perf_counter()
        print("rep %d: scan %.1f ms, align_batch %.1f ms (%d alignments, %d accepted)" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, n_aln, sum(len(x) for x in r)))
    ctx.close()


if __name__ == "__main__":
    main()
