"""scripts/scan_classes.py -- scan throughput per query length (development aid): run once as is and once with
B200_SCAN_COARSE=1 to compare capacity-class lists."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mmseqs2_b200 import Context, SubMatrix, synth  # noqa: E402


def main():
    mat, pb = bench.load_matrix()
    sm = SubMatrix(mat, pb)
    bg = synth.background(pb)
    rng = np.random.default_rng(3)
    res, off = synth.random_seqs(rng, int(os.environ.get("DBSEQS", "200000")), bg, mean=300, sigma=0.6, lo=30, hi=5000)
    ctx = Context(0)
    ctx.load_db(res, off, 21)
    nres = int(off[-1])
    print("coarse" if os.environ.get("B200_SCAN_COARSE") else "fine", "class list; DB residues", nres)
    for L in [int(x) for x in os.environ.get("LENS", "200,260,300,330,350,370,400,440,480,600,700,900,1100,1500").split(",")]:
        qs = [synth.random_seqs(rng, 1, bg, mean=L, sigma=0, lo=L, hi=L, normal=True)[0] for _ in range(16)]
        job = ctx.scan_job([sm.ssw_query(q) for q in qs], 15, 300)
        job.run(); ctx.sync()
        ctx.event_record(0)
        for _ in range(3):
            job.run()
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms(0, 1) / 3
        print("L=%4d  %.2f ms  %.0f GCUPS" % (L, ms, 16 * L * nres / 1e9 / (ms / 1e3)))
        job.close()
    ctx.close()


if __name__ == "__main__":
    main()
