"""scripts/align_step_perf.py -- phase times of b200_align_batch on the config[2]-shaped lists (run with B200_TRACE=1)."""
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
    nq = int(os.environ.get("SWQ", "512")); nt = int(os.environ.get("SWT", "256"))
    mat, pb = bench.load_matrix()
    sm = SubMatrix(mat, pb)
    sq, std, sto, spairs = bench.make_sw_workload(nq, nt)
    ctx = Context(0)
    ctx.load_db(std, sto, 21)
    order = np.argsort(spairs[:, 0], kind="stable")
    bounds = np.searchsorted(spairs[order, 0], np.arange(len(sq) + 1))
    lists = [spairs[order[bounds[i]:bounds[i + 1]], 1] for i in range(len(sq))]
    evp = al.EvalueParams.defaults("blosum62.out", 11, 1, int(sto[-1]))
    for mode in (2, 1):
        par = al.AlignParams(sw_mode=mode, eval_thr=1e-3)
        al.align_batch(ctx, sm, sq[:8], lists[:8], par, evp)
        for rep in range(2):   # the first full-size call grows the device buffer pool
            t0 = time.perf_counter()
            res, pool, n_aln = al.align_batch(ctx, sm, sq, lists, par, evp)
            dt = time.perf_counter() - t0
        print("mode %d: %d alignments, %d accepted, %.1f ms, %.0f alignments/s" % (mode, n_aln, sum(len(r) for r in res), dt * 1e3, n_aln / dt))
    ctx.close()


if __name__ == "__main__":
    main()
