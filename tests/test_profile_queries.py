"""Profile (PSSM) queries -- SURVEY 8f row 4: the HMM_PROFILE branches of ssw_init / createProfile feed the same kernels an
[A][L] int8 table instead of substitution-matrix rows.  Fixtures: tests/golden/profile_v1.npz, written by the reference's own
profile code paths (tests/golden/make_profile_golden.py)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pg():
    return np.load(os.path.join(ROOT, "tests", "golden", "profile_v1.npz"))


def test_pssm_builder(pg, submat):
    for k in range(int(pg["n"])):
        p = pg["pssm%d" % k]
        prof = submat.pssm_query(p)
        assert prof.profile.shape == (21, p.shape[1])
        assert np.array_equal(prof.profile[:20], p) and not prof.profile[20].any()      # X is the neutral state
        assert prof.bias == max(0, -int(p.min()))


@pytest.mark.gpu
def test_profile_scan_and_alignment(pg, ctx, submat):
    n = len(pg["toff"]) - 1
    ctx.load_db(pg["tdata"], pg["toff"].astype(np.uint64), 21)
    nq = int(pg["n"])
    profs = [submat.pssm_query(pg["pssm%d" % k]) for k in range(nq)]
    cons = [pg["cons%d" % k] for k in range(nq)]
    _, _, dense = ctx.ungapped_scan(profs, want_dense=True)
    pairs = np.array([(qi, t) for qi in range(nq) for t in range(n)], np.uint32)
    aln = ctx.sw_align(profs, pairs)
    out, bts = ctx.sw_backtrace(profs, cons, pairs, aln)
    checked = 0
    for k in range(nq):
        assert np.array_equal(dense[k].astype(np.int32), pg["ungapped%d" % k]), k
        sl = slice(k * n, (k + 1) * n)
        got = np.stack([aln[f][sl] for f in ("score", "qstart", "qend", "dbstart", "dbend", "word")], 1)
        exp = pg["align%d" % k]
        assert np.array_equal(got, exp), (k, np.nonzero((got != exp).any(1))[0][:5])
        for t in range(n):
            i = k * n + t
            if aln["dbend"][i] == -1:
                continue
            assert out["ok"][i] == 1 and bts[i] == pg["bt%d" % k][t].decode() and out["identical"][i] == pg["ident%d" % k][t], (k, t)
            checked += 1
    assert checked > 1500
    assert int(sum(pg["align%d" % k][:, 5].sum() for k in range(nq))) > 100        # word-mode pairs are part of the fixture


@pytest.mark.gpu
def test_profile_diagonal_scorer(pg, ctx, submat):
    ctx.load_db(pg["tdata"], pg["toff"].astype(np.uint64), 21)
    for k in range(int(pg["n"])):
        prof = submat.pssm_query(pg["pssm%d" % k])
        cnt, raw = ctx.diag_score(prof, pg["diag_ids%d" % k], pg["diag_dg%d" % k], want_raw=True)
        assert np.array_equal(cnt, pg["diag_counts%d" % k]), k
        assert np.array_equal(raw, pg["diag_raw%d" % k]), k
