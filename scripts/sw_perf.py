"""scripts/sw_perf.py -- quick SW-only timing (development aid): packed score kernel, int32 score+end kernel, host path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mmseqs2_b200 import Context, SubMatrix  # noqa: E402


def main():
    nq = int(os.environ.get("SWQ", "512")); nt = int(os.environ.get("SWT", "256"))
    mat, pb = bench.load_matrix()
    sm = SubMatrix(mat, pb)
    t0 = time.time()
    sq, std, sto, spairs = bench.make_sw_workload(nq, nt)
    ctx = Context(0)
    ctx.load_db(std, sto, 21)
    profs = [sm.ssw_query(q) for q in sq]
    print("setup %.1fs, %d pairs" % (time.time() - t0, len(spairs)))
    job = ctx.sw_score_job(profs, spairs)
    job.run(); ctx.sync()
    ctx.event_record(0)
    for _ in range(5):
        job.run()
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1) / 5
    cells = job.cells
    print("sw16 warps=%s: %.3f ms  %.1f GCUPS  frac=%.3f" % (os.environ.get("B200_SW16_WARPS", "8"), ms, cells / 1e9 / (ms / 1e3),
                                                         cells * 3.25 / (ms / 1e3) / (64 * 148 * 1.965e9)))
    job.close()
    for rep in range(3):
        t0 = time.perf_counter()
        ctx.sw_score(profs, spairs)
        dt = time.perf_counter() - t0
    print("e2e sw_score: %.1f ms  %.1f GCUPS" % (dt * 1e3, cells / 1e9 / dt))
    for rep in range(3):
        t0 = time.perf_counter()
        ends = ctx.sw_score_endpos(profs, spairs)
        dt = time.perf_counter() - t0
    print("e2e sw_score_endpos (packed score + packed FIND): %.1f ms  %.1f GCUPS" % (dt * 1e3, cells / 1e9 / dt))
    for rep in range(2):
        t0 = time.perf_counter()
        aln = ctx.sw_align(profs, spairs)
        dt = time.perf_counter() - t0
    print("e2e sw_align (score + end + start, all pairs): %.1f ms  %.1f GCUPS" % (dt * 1e3, cells / 1e9 / dt))
    sjob = ctx.sw_job(profs, spairs)
    sjob.run(); ctx.sync()
    ctx.event_record(2); sjob.run(); ctx.event_record(3)
    ms2 = ctx.event_elapsed_ms(2, 3)
    print("sw32 score+end: %.3f ms  %.1f GCUPS" % (ms2, cells / 1e9 / (ms2 / 1e3)))
    e32 = sjob.fetch()
    assert np.array_equal(e32, ends), "packed FIND path differs from the int32 kernel"
    sjob.close()
    ctx.close()


if __name__ == "__main__":
    main()
