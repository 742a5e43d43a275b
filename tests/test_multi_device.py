"""GPU (-m gpu): the multi-device handle (include/b200_multi.h).  On a one-GPU box the handle is built over two contexts of the same
device -- the sharding and merge logic is identical; with two or more GPUs visible the real devices are used."""
import numpy as np
import pytest

from mmseqs2_b200 import synth

pytestmark = pytest.mark.gpu


def _devices():
    from mmseqs2_b200 import load_library
    n = load_library().b200_device_count()
    return list(range(n)) if n >= 2 else [0, 0, 0]


def _workload(blosum, n_db=30000, nq=11):
    rng = np.random.default_rng(99)
    bg = synth.background(blosum[1])
    res, off = synth.random_seqs(rng, n_db, bg, mean=250, sigma=0.6, lo=20, hi=3000)
    qs = synth.split(*synth.random_seqs(rng, nq, bg, mean=330, sigma=90, lo=60, hi=900, normal=True))
    synth.plant_homologs(rng, res, off, qs, bg, frac=0.05)
    return res, off, qs


def test_scan_is_independent_of_the_split(ctx, submat, blosum):
    from mmseqs2_b200 import MultiContext
    res, off, qs = _workload(blosum)
    profs = [submat.ssw_query(q) for q in qs]
    ctx.load_db(res, off, 21)
    h1, n1, _ = ctx.ungapped_scan(profs, 15, 120)
    for shard_targets in (False, True):
        m = MultiContext(_devices())
        assert m.size >= 2
        m.load_db(res, off, 21, shard_targets=shard_targets)
        h2, n2 = m.ungapped_scan(profs, 15, 120)
        m.close()
        assert np.array_equal(n1, n2), shard_targets
        for i in range(len(qs)):
            assert np.array_equal(h1[i][:n1[i]], h2[i][:n2[i]]), (shard_targets, i)
    # ties at the cut-off: a low threshold makes every list full, so the merged tail depends on the (score, global id) order
    h1, n1, _ = ctx.ungapped_scan(profs, 3, 50)
    m = MultiContext(_devices())
    m.load_db(res, off, 21, shard_targets=True)
    h2, n2 = m.ungapped_scan(profs, 3, 50)
    m.close()
    assert (n1 == 50).all() and np.array_equal(n1, n2) and np.array_equal(h1, h2)


def test_align_batch_is_independent_of_the_split(ctx, submat, blosum):
    from mmseqs2_b200 import MultiContext, alignment as al
    res, off, qs = _workload(blosum, n_db=8000, nq=9)
    profs = [submat.ssw_query(q) for q in qs]
    ctx.load_db(res, off, 21)
    h, n, _ = ctx.ungapped_scan(profs, 15, 200)
    lists = [h[i]["id"][:int(n[i])] for i in range(len(qs))]
    evp = al.EvalueParams.defaults("blosum62.out", 11, 1, int(off[-1]))
    par = al.AlignParams(sw_mode=al.SCORE_COV_SEQID, eval_thr=10.0)
    r1, p1, a1 = al.align_batch(ctx, submat, qs, lists, par, evp)
    m = MultiContext(_devices())
    m.load_db(res, off, 21)
    r2, p2, a2 = al.align_batch(m, submat, qs, lists, par, evp)
    m.close()
    assert a1 == a2 and sum(len(r) for r in r1) > 20
    for i in range(len(qs)):
        assert al.records(r1[i], p1, True, True) == al.records(r2[i], p2, True, True), i
