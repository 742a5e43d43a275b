"""GPU (-m gpu): the CUDA path through the C ABI against (1) the committed reference outputs, (2) the oracle on seeded
inputs, (3) size-independent properties at larger sizes.  Bit-exact everywhere (integer path)."""
import numpy as np
import pytest

from mmseqs2_b200 import synth
from oracle.pyoracle import pack_targets

pytestmark = pytest.mark.gpu


def _queries(golden):
    return [golden["q%d" % i] for i in range(int(golden["nq"]))]


def _expected_hits(dense, thr, k):
    ids = np.nonzero(dense > thr)[0]
    order = np.lexsort((ids, -dense[ids].astype(np.int64)))
    ids = ids[order][:k]
    return ids.astype(np.uint32), dense[ids].astype(np.int32)


@pytest.fixture()
def golden_db(ctx, golden):
    ctx.load_db(golden["tdata"], golden["toff"].astype(np.uint64), 21)
    return golden["tdata"], golden["toff"]


def test_scan_matches_reference_outputs(ctx, golden, golden_db, submat):
    qs = _queries(golden)
    for cbf in (0, 1):
        profs = [submat.ssw_query(q, comp_bias=bool(cbf)) for q in qs]
        hits, n_hits, dense = ctx.ungapped_scan(profs, min_score_excl=15, max_hits=50, want_dense=True)
        for i in range(len(qs)):
            exp = golden["q%d_cb%d_ungapped" % (i, cbf)]
            assert np.array_equal(dense[i].astype(np.int32), exp), (i, cbf)
            eid, esc = _expected_hits(exp, 15, 50)
            assert n_hits[i] == len(eid)
            assert np.array_equal(hits[i]["id"][:len(eid)], eid) and np.array_equal(hits[i]["score"][:len(eid)], esc)


def test_sw_matches_reference_outputs(ctx, golden, golden_db, submat):
    qs = _queries(golden)
    n = len(golden["toff"]) - 1
    pairs = np.array([(qi, t) for qi in range(len(qs)) for t in range(n)], np.uint32)
    for cbf in (0, 1):
        profs = [submat.ssw_query(q, comp_bias=bool(cbf)) for q in qs]
        ends = ctx.sw_score_endpos(profs, pairs)
        aln = ctx.sw_align(profs, pairs)
        for qi in range(len(qs)):
            e = golden["q%d_cb%d_endpos" % (qi, cbf)]
            a = golden["q%d_cb%d_align" % (qi, cbf)]
            sl = slice(qi * n, (qi + 1) * n)
            got_e = np.stack([ends["score"][sl], ends["qend"][sl], ends["dbend"][sl], ends["word"][sl]], 1)
            assert np.array_equal(got_e, e), (qi, cbf, np.nonzero((got_e != e).any(1))[0][:5])
            got_a = np.stack([aln[f][sl] for f in ("score", "qstart", "qend", "dbstart", "dbend", "word")], 1)
            assert np.array_equal(got_a, a), (qi, cbf, np.nonzero((got_a != a).any(1))[0][:5])


def test_packed_score_matches_reference_outputs(ctx, golden, golden_db, submat):
    """b200_sw_score (int16x2, two targets per warp) == score1 of alignScoreEndPos for every pair of the fixture"""
    qs = _queries(golden)
    n = len(golden["toff"]) - 1
    pairs = np.array([(qi, t) for qi in range(len(qs)) for t in range(n)], np.uint32)
    for cbf in (0, 1):
        profs = [submat.ssw_query(q, comp_bias=bool(cbf)) for q in qs]
        sc = ctx.sw_score(profs, pairs)
        for qi in range(len(qs)):
            exp = golden["q%d_cb%d_endpos" % (qi, cbf)][:, 0]
            got = sc[qi * n:(qi + 1) * n]
            assert np.array_equal(got, exp), (qi, cbf, np.nonzero(got != exp)[0][:5], got[got != exp][:5], exp[got != exp][:5])
    # odd pair counts, single pairs, gap variants (go < ge takes the int32 kernel)
    profs = [submat.ssw_query(q) for q in qs]
    for go, ge in ((11, 1), (5, 2), (2, 3)):
        sub = pairs[(pairs[:, 0] % 3 == 1)][:-1:7]
        got = ctx.sw_score(profs, sub, go=go, ge=ge)
        exp = ctx.sw_score_endpos(profs, sub, go=go, ge=ge)["score"]
        assert np.array_equal(got, exp), (go, ge)
    one = ctx.sw_score(profs, pairs[5:6])
    assert one[0] == golden["q0_cb1_endpos"][5, 0]


def test_backtrace_matches_reference_outputs(ctx, golden, golden_db, submat):
    """A6: b200_sw_backtrace == the reference's backtrace string and identity count (alignment mode 2) for the fixture"""
    qs = _queries(golden)
    n = len(golden["toff"]) - 1
    pairs = np.array([(qi, t) for qi in range(len(qs)) for t in range(n)], np.uint32)
    profs = [submat.ssw_query(q) for q in qs]
    aln = ctx.sw_align(profs, pairs)
    out, bts = ctx.sw_backtrace(profs, qs, pairs, aln)
    checked = 0
    for qi in range(len(qs)):
        exp_bt, exp_id = golden["q%d_cb1_bt" % qi], golden["q%d_cb1_ident" % qi]
        for t in range(n):
            k = qi * n + t
            if aln["dbend"][k] == -1:
                assert out["n_cigar"][k] == 0
                continue
            assert out["ok"][k] == 1 and bts[k] == str(exp_bt[t]) and out["identical"][k] == exp_id[t], (qi, t)
            assert out["bt_len"][k] == len(bts[k])
            checked += 1
    assert checked > 2000


def test_backtrace_gap_penalties_and_wide_bands(ctx, oracle, submat, blosum):
    """A6 against the oracle for gap costs other than 11/1 (go > ge, go == ge, go < ge: the warp kernel's F chain uses min(go, ge))
    and for alignments whose band has to double many times (long indels; bands beyond the shared-memory rows)"""
    rng = np.random.default_rng(4242)
    bg = synth.background(blosum[1])
    q = synth.random_seqs(rng, 1, bg, mean=600, sigma=0, lo=600, hi=600, normal=True)[0]
    tg = [synth.mutate(rng, q, bg, s, i) for s, i in ((0.1, 0.02), (0.3, 0.05), (0.2, 0.0), (0.05, 0.1), (0.4, 0.03))]
    tg.append(np.concatenate([q[:200], q[420:]]))                                    # one 220-residue deletion: band >= 220
    tg.append(np.concatenate([q[:100], synth.random_seqs(rng, 1, bg, mean=300, sigma=0, lo=300, hi=300, normal=True)[0], q[100:]]))
    tg.append(q[50:550].copy())
    tg += [synth.mutate(rng, q[a:a + 250], bg, 0.25, 0.04) for a in (0, 100, 350)]
    td, to = pack_targets(tg)
    ctx.load_db(td, to.astype(np.uint64), 21)
    prof = submat.ssw_query(q)
    cb, bias = oracle.query_cb(q, True)
    pairs = np.array([(0, t) for t in range(len(tg))], np.uint32)
    checked = 0
    for go, ge in ((11, 1), (5, 2), (3, 3), (2, 4), (20, 1)):
        aln = ctx.sw_align([prof], pairs, go=go, ge=ge)
        out, bts = ctx.sw_backtrace([prof], [q], pairs, aln, go=go, ge=ge)
        for k in range(len(tg)):
            if aln["dbend"][k] == -1:
                continue
            row = [int(aln[f][k]) for f in ("score", "qstart", "qend", "dbstart", "dbend")]
            exp_bt, exp_id = oracle.backtrace(q, cb, tg[k], row, go, ge)
            assert exp_bt is not None and out["ok"][k] == 1, (go, ge, k)
            assert bts[k] == exp_bt and out["identical"][k] == exp_id, (go, ge, k)
            checked += 1
    assert checked >= 40


def test_diag_matches_reference_outputs(ctx, golden, golden_db, submat):
    for qi, q in enumerate(_queries(golden)):
        ids, dg = golden["q%d_diag_ids" % qi], golden["q%d_diag_dg" % qi]
        for cbf in (0, 1):
            prof = submat.diag_query(q, golden["q%d_compbias" % qi] if cbf else None)
            cnt, raw = ctx.diag_score(prof, ids, dg, want_raw=True)
            assert np.array_equal(cnt, golden["q%d_cb%d_diag_counts" % (qi, cbf)]), (qi, cbf)
            assert np.array_equal(raw, golden["q%d_cb%d_diag_raw" % (qi, cbf)]), (qi, cbf)
    # entries with a non-zero count on input are left alone (UngappedAlignment.cpp:327-329)
    q = _queries(golden)[5]
    ids, dg = golden["q5_diag_ids"], golden["q5_diag_dg"]
    pre = np.zeros(len(ids), np.uint8)
    pre[::3] = 7
    cnt, _ = ctx.diag_score(submat.diag_query(q, None), ids, dg, counts=pre)
    exp = golden["q5_cb0_diag_counts"].copy()
    exp[::3] = 7
    assert np.array_equal(cnt, exp)


@pytest.mark.parametrize("seed", [5, 6])
def test_seeded_random_vs_oracle(ctx, oracle, submat, blosum, seed):
    rng = np.random.default_rng(seed)
    bg = synth.background(blosum[1])
    res, off = synth.random_seqs(rng, 600, bg, mean=220, sigma=0.8, lo=1, hi=3000)
    qlens = [3, 40, 127, 128, 129, 300, 511, 512, 513, 900, 2047]
    qs = [synth.random_seqs(rng, 1, bg, mean=L, sigma=0, lo=L, hi=L, normal=True)[0] for L in qlens]
    synth.plant_homologs(rng, res, off, [q for q in qs if len(q) > 30], bg, frac=0.4, subst=0.2, indel=0.03)
    td, to = res, off.astype(np.int64)
    ctx.load_db(td, off, 21)
    profs = [submat.ssw_query(q) for q in qs]
    _, _, dense = ctx.ungapped_scan(profs, want_dense=True)
    n = len(off) - 1
    pairs = np.array([(qi, t) for qi in range(len(qs)) for t in rng.choice(n, 150, replace=False)], np.uint32)
    aln = ctx.sw_align(profs, pairs)
    assert np.array_equal(ctx.sw_score(profs, pairs), aln["score"])
    for qi, q in enumerate(qs):
        cb, bias = oracle.query_cb(q, True)
        assert np.array_equal(dense[qi].astype(np.int32), oracle.ungapped(q, cb, bias, td, to)), qi
        sel = np.nonzero(pairs[:, 0] == qi)[0]
        tg = [td[int(to[t]):int(to[t + 1])] for t in pairs[sel, 1]]
        sd, so = pack_targets(tg)
        exp = oracle.sw_align(q, cb, bias, sd, so)
        got = np.stack([aln[f][sel] for f in ("score", "qstart", "qend", "dbstart", "dbend", "word")], 1)
        assert np.array_equal(got, exp), (qi, np.nonzero((got != exp).any(1))[0][:5])


def test_long_query_tiles_and_gap_variants(ctx, oracle, submat, blosum):
    rng = np.random.default_rng(77)
    bg = synth.background(blosum[1])
    q = synth.random_seqs(rng, 1, bg, mean=2600, sigma=0, lo=2600, hi=2600, normal=True)[0]
    tg = [synth.mutate(rng, q, bg, 0.15, 0.02), synth.mutate(rng, q[300:2100], bg, 0.4, 0.05),
          synth.random_seqs(rng, 1, bg, mean=3000, sigma=0, lo=3000, hi=3000, normal=True)[0], q.copy(), q[:5].copy()]
    td, to = pack_targets(tg)
    ctx.load_db(td, to.astype(np.uint64), 21)
    prof = submat.ssw_query(q)
    cb, bias = oracle.query_cb(q, True)
    pairs = np.array([(0, t) for t in range(len(tg))], np.uint32)
    for go, ge in ((11, 1), (5, 2), (20, 3)):
        assert np.array_equal(ctx.sw_score([prof], pairs, go=go, ge=ge), oracle.sw_align(q, cb, bias, td, to, go, ge)[:, 0])
        aln = ctx.sw_align([prof], pairs, go=go, ge=ge)
        got = np.stack([aln[f] for f in ("score", "qstart", "qend", "dbstart", "dbend", "word")], 1)
        assert np.array_equal(got, oracle.sw_align(q, cb, bias, td, to, go, ge)), (go, ge)


def test_scan_every_capacity_class(ctx, oracle, submat, blosum):
    """one query at each capacity boundary of the scan kernel's (G,K) classes: qlen = capacity-1 (last row of a class) and
    qlen = capacity (first length of the next class) for 32-row steps up to 512, 64-row steps up to 1024, 128-row steps beyond"""
    rng = np.random.default_rng(777)
    bg = synth.background(blosum[1])
    caps = list(range(64, 513, 32)) + list(range(576, 1025, 64)) + list(range(1152, 2049, 128))
    lens = sorted(set([c - 1 for c in caps] + [c for c in caps[:-1]]))
    res, off = synth.random_seqs(rng, 120, bg, mean=300, sigma=0.8, lo=1, hi=2500)
    qs = [synth.random_seqs(rng, 1, bg, mean=L, sigma=0, lo=L, hi=L, normal=True)[0] for L in lens]
    synth.plant_homologs(rng, res, off, qs, bg, frac=0.6, subst=0.15, indel=0.01)
    seqs = synth.split(res, off)
    for k in range(0, len(qs), 5):                 # near-identical targets: saturation at 255 - bias in the top rows of a class
        seqs[k % len(seqs)] = qs[k].copy()
    td, to = pack_targets(seqs)
    ctx.load_db(td, to.astype(np.uint64), 21)
    profs = [submat.ssw_query(q) for q in qs]
    _, _, dense = ctx.ungapped_scan(profs, want_dense=True)
    for qi, q in enumerate(qs):
        cb, bias = oracle.query_cb(q, True)
        exp = oracle.ungapped(q, cb, bias, td, to)
        assert np.array_equal(dense[qi].astype(np.int32), exp), (len(q), np.nonzero(dense[qi] != exp)[0][:5])


def test_scan_long_queries_tiled(ctx, oracle, submat, blosum):
    """queries beyond one 2048-row tile (2048, 2049, 4095, 4096, 5000 rows) and a 1-residue query"""
    rng = np.random.default_rng(2048)
    bg = synth.background(blosum[1])
    res, off = synth.random_seqs(rng, 150, bg, mean=400, sigma=0.9, lo=1, hi=6000)
    qs = [synth.random_seqs(rng, 1, bg, mean=L, sigma=0, lo=L, hi=L, normal=True)[0] for L in (2048, 2049, 4095, 4096, 5000, 1)]
    synth.plant_homologs(rng, res, off, qs[:5], bg, frac=0.5, subst=0.1, indel=0.01)
    seqs = synth.split(res, off)
    seqs[0] = qs[2].copy(); seqs[1] = qs[4][100:4900].copy()     # long identical diagonals crossing tile boundaries
    td, to = pack_targets(seqs)
    ctx.load_db(td, to.astype(np.uint64), 21)
    profs = [submat.ssw_query(q) for q in qs]
    _, _, dense = ctx.ungapped_scan(profs, want_dense=True)
    for qi, q in enumerate(qs):
        cb, bias = oracle.query_cb(q, True)
        exp = oracle.ungapped(q, cb, bias, td, to)
        assert np.array_equal(dense[qi].astype(np.int32), exp), (qi, np.nonzero(dense[qi] != exp)[0][:5])
    assert dense[2][0] > 200 and dense[4][1] > 200


def test_gate_and_passthrough(ctx, oracle, submat, blosum, golden, golden_db):
    qs = _queries(golden)[4:7]
    profs = [submat.ssw_query(q) for q in qs]
    n = len(golden["toff"]) - 1
    pairs = np.array([(qi, t) for qi in range(len(qs)) for t in range(0, n, 3)], np.uint32)
    gate = (np.arange(len(pairs)) % 2).astype(np.uint8)
    full = ctx.sw_align(profs, pairs)
    gated = ctx.sw_align(profs, pairs, gate=gate)
    for f in ("score", "qend", "dbend", "word"):
        assert np.array_equal(full[f], gated[f])
    on = gate == 1
    assert np.array_equal(full["qstart"][on], gated["qstart"][on]) and np.array_equal(full["dbstart"][on], gated["dbstart"][on])
    assert (gated["qstart"][~on] == -1).all() and (gated["dbstart"][~on] == -1).all()


def test_properties_at_scale(ctx, submat, blosum):
    """size-independent checks on a DB too big for the scalar oracle: permutation equivariance, run-to-run determinism,
    symmetry of the gapped score without composition bias, self-alignment bounds."""
    rng = np.random.default_rng(4242)
    bg = synth.background(blosum[1])
    res, off = synth.random_seqs(rng, 50000, bg, mean=300, sigma=0.6, lo=30, hi=5000)
    qres, qoff = synth.random_seqs(rng, 6, bg, mean=350, sigma=35, lo=200, hi=500, normal=True)
    qs = synth.split(qres, qoff)
    synth.plant_homologs(rng, res, off, qs, bg, frac=0.01)
    profs = [submat.ssw_query(q) for q in qs]
    ctx.load_db(res, off, 21)
    h1, n1, d1 = ctx.ungapped_scan(profs, want_dense=True)
    job = ctx.scan_job(profs)
    job.run(); job.run()
    h2, n2, d2 = job.fetch(want_dense=True)
    job.close()
    assert np.array_equal(d1, d2) and np.array_equal(n1, n2) and np.array_equal(h1, h2)
    assert int(d1.astype(np.int64).sum()) > 0
    # permute the DB: dense scores must follow the permutation
    n = len(off) - 1
    perm = rng.permutation(n)
    seqs = synth.split(res, off)
    pres, poff = pack_targets([seqs[i] for i in perm])
    ctx.load_db(pres, poff.astype(np.uint64), 21)
    _, _, d3 = ctx.ungapped_scan(profs, want_dense=True)
    assert np.array_equal(d3, d1[:, perm])
    # gapped score symmetry score(q,t) == score(t,q) for a symmetric matrix without composition bias
    sub = rng.choice(n, 300, replace=False)
    tq = [submat.ssw_query(seqs[perm[i]], comp_bias=False) for i in sub[:20]]
    qp = [submat.ssw_query(q, comp_bias=False) for q in qs]
    fwd = ctx.sw_score_endpos(qp, np.array([(a, sub[b]) for a in range(len(qs)) for b in range(20)], np.uint32))
    tdb, tdo = pack_targets(qs)
    ctx.load_db(tdb, tdo.astype(np.uint64), 21)
    bwd = ctx.sw_score_endpos(tq, np.array([(b, a) for a in range(len(qs)) for b in range(20)], np.uint32))
    assert np.array_equal(fwd["score"], bwd["score"])
    # self alignment: score == sum of diagonal scores, end = last residue, start = 0
    selfp = ctx.sw_align(qp, np.array([(a, a) for a in range(len(qs))], np.uint32))
    for a, q in enumerate(qs):
        assert selfp["score"][a] == int(blosum[0][q, q].astype(np.int64).sum())
        assert selfp["qend"][a] == len(q) - 1 and selfp["dbend"][a] == len(q) - 1
        assert selfp["qstart"][a] == 0 and selfp["dbstart"][a] == 0


def test_shared_context_from_many_host_threads(ctx, golden, golden_db, submat):
    """the reference keeps one operator object per OpenMP thread; here they share one ctx (internal mutex)"""
    import threading
    qs = _queries(golden)[3:9]
    profs = [submat.ssw_query(q) for q in qs]
    n = len(golden["toff"]) - 1
    pairs = np.array([(qi, t) for qi in range(len(qs)) for t in range(0, n, 2)], np.uint32)
    exp_sc = ctx.sw_score(profs, pairs)
    exp_hits, exp_n, _ = ctx.ungapped_scan(profs, max_hits=40)
    errors = []

    def worker(k):
        try:
            for _ in range(4):
                if k % 2:
                    assert np.array_equal(ctx.sw_score(profs, pairs), exp_sc)
                else:
                    h, nh, _ = ctx.ungapped_scan(profs, max_hits=40)
                    assert np.array_equal(h, exp_hits) and np.array_equal(nh, exp_n)
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_empty_and_tiny_sequences(ctx, oracle, submat, blosum):
    """zero-length targets, 1-residue targets/queries, a DB of a single sequence"""
    rng = np.random.default_rng(3)
    bg = synth.background(blosum[1])
    seqs = [np.zeros(0, np.uint8), np.array([5], np.uint8), synth.random_seqs(rng, 1, bg, mean=40, sigma=0, lo=40, hi=40, normal=True)[0],
            np.zeros(0, np.uint8), np.array([19, 19], np.uint8)]
    td, to = pack_targets(seqs)
    ctx.load_db(td, to.astype(np.uint64), 21)
    qs = [np.array([5], np.uint8), seqs[2][5:30].copy(), np.array([19, 19, 19], np.uint8)]
    profs = [submat.ssw_query(q) for q in qs]
    _, _, dense = ctx.ungapped_scan(profs, want_dense=True)
    pairs = np.array([(qi, t) for qi in range(len(qs)) for t in range(len(seqs))], np.uint32)
    aln = ctx.sw_align(profs, pairs)
    sc = ctx.sw_score(profs, pairs)
    for qi, q in enumerate(qs):
        cb, bias = oracle.query_cb(q, True)
        nz = [k for k in range(len(seqs)) if len(seqs[k]) > 0]
        sd, so = pack_targets([seqs[k] for k in nz])
        assert np.array_equal(dense[qi][nz].astype(np.int32), oracle.ungapped(q, cb, bias, sd, so))
        exp = oracle.sw_align(q, cb, bias, sd, so)
        got = np.stack([aln[f][qi * len(seqs):(qi + 1) * len(seqs)] for f in ("score", "qstart", "qend", "dbstart", "dbend", "word")], 1)
        assert np.array_equal(got[nz], exp), qi
        for k in range(len(seqs)):
            if len(seqs[k]) == 0:   # nothing to align: score 0, "no residue aligned"
                assert dense[qi][k] == 0 and got[k, 0] == 0 and got[k, 4] == -1 and sc[qi * len(seqs) + k] == 0
    one = pack_targets([seqs[2]])
    ctx.load_db(one[0], one[1].astype(np.uint64), 21)
    h, nh, _ = ctx.ungapped_scan([profs[1]], min_score_excl=0, max_hits=5)
    assert nh[0] == 1 and h[0]["id"][0] == 0


def test_error_paths(ctx, submat, blosum):
    from mmseqs2_b200 import B200Error
    rng = np.random.default_rng(1)
    bg = synth.background(blosum[1])
    res, off = synth.random_seqs(rng, 10, bg, mean=50, sigma=0.1, lo=10, hi=100)
    ctx.load_db(res, off, 21)
    q = submat.ssw_query(res[:30])
    with pytest.raises(B200Error):
        ctx.sw_score_endpos([q], np.array([(0, 10)], np.uint32))   # target id out of range
    with pytest.raises(B200Error):
        ctx.sw_score_endpos([q], np.array([(1, 0)], np.uint32))    # query index out of range
    bad = res.copy(); bad[3] = 40
    with pytest.raises(B200Error):
        ctx.load_db(bad, off, 21)                                   # masked (+32) residue must be stripped first
    ctx.load_db(res, off, 21)
    out = ctx.sw_score_endpos([q], np.zeros((0, 2), np.uint32))     # empty batch is fine
    assert len(out) == 0


def test_scan_at_scale_vs_oracle(ctx, oracle, submat, blosum):
    """the scan where it is benchmarked: a 200 000-sequence DB (every target-length regime, persistent CTAs with dynamic units,
    several capacity classes in one call) against the multi-threaded C restatement -- every score and the hit lists."""
    import os
    rng = np.random.default_rng(20260923)
    bg = synth.background(blosum[1])
    res, off = synth.random_seqs(rng, 200000, bg, mean=300, sigma=0.6, lo=30, hi=5000)
    qs = [synth.random_seqs(rng, 1, bg, mean=L, sigma=0, lo=L, hi=L, normal=True)[0] for L in (207, 351, 352, 489)]
    synth.plant_homologs(rng, res, off, qs, bg, frac=0.01)
    synth.plant_homologs(rng, res, off, qs, bg, frac=0.002, subst=0.05)     # near-identical copies: saturating scores
    profs = [submat.ssw_query(q) for q in qs]
    ctx.load_db(res, off, 21)
    hits, n_hits, dense = ctx.ungapped_scan(profs, min_score_excl=15, max_hits=300, want_dense=True)
    off64 = off.astype(np.int64)
    nthreads = len(os.sched_getaffinity(0))
    saturated = 0
    for i, q in enumerate(qs):
        cb, bias = oracle.query_cb(q, True)
        exp = oracle.ungapped(q, cb, bias, res, off64, nthreads=nthreads, fast=True)
        assert np.array_equal(dense[i].astype(np.int32), exp), (i, np.nonzero(dense[i] != exp)[0][:5])
        eid, esc = _expected_hits(exp, 15, 300)
        assert n_hits[i] == len(eid)
        assert np.array_equal(hits[i]["id"][:len(eid)], eid) and np.array_equal(hits[i]["score"][:len(eid)], esc)
        saturated += int((exp == 255 - bias).sum())
    assert saturated > 0


def test_backtrace_unreachable_score_terminates(ctx, golden, golden_db, submat):
    """a claimed score that no path inside the rectangle reaches (caller error, or a score capped at 32767): the band search stops
    once the band covers the rectangle and the task comes back with ok == 0 -- it must neither run away nor write out of bounds"""
    qs = _queries(golden)
    n = len(golden["toff"]) - 1
    pairs = np.array([(0, t) for t in range(n)], np.uint32)
    profs = [submat.ssw_query(q) for q in qs[:1]]
    aln = ctx.sw_align(profs, pairs)
    good = np.nonzero(aln["dbend"] != -1)[0]
    assert len(good) > 10
    bad = aln.copy()
    bad["score"][good[::2]] += 500
    out, bts = ctx.sw_backtrace(profs, qs[:1], pairs, bad)
    assert (out["ok"][good[::2]] == 0).all() and (out["n_cigar"][good[::2]] == 0).all()
    ref_out, ref_bts = ctx.sw_backtrace(profs, qs[:1], pairs, aln)
    for k in good[1::2]:
        assert out["ok"][k] == 1 and bts[k] == ref_bts[k]
