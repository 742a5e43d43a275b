"""CPU: the C restatement (oracle/oracle.c) against the committed reference outputs (tests/golden/hotpath_v1.npz,
generated from the reference's own code by tests/golden/make_golden.py)."""
import numpy as np


def _queries(golden):
    return [golden["q%d" % i] for i in range(int(golden["nq"]))]


def test_comp_bias_bits(golden, oracle):
    for i, q in enumerate(_queries(golden)):
        f = oracle.comp_bias(q)
        assert np.array_equal(f.view(np.uint32), golden["q%d_compbias" % i].view(np.uint32))


def test_ungapped_scan_scores(golden, oracle):
    td, to = golden["tdata"], golden["toff"]
    for i, q in enumerate(_queries(golden)):
        for cbf in (0, 1):
            cb, bias = oracle.query_cb(q, bool(cbf))
            got = oracle.ungapped(q, cb, bias, td, to)
            assert np.array_equal(got, golden["q%d_cb%d_ungapped" % (i, cbf)]), (i, cbf)


def test_sw_score_endpos(golden, oracle):
    td, to = golden["tdata"], golden["toff"]
    words = 0
    for i, q in enumerate(_queries(golden)):
        for cbf in (0, 1):
            cb, bias = oracle.query_cb(q, bool(cbf))
            got = oracle.sw_score_endpos(q, cb, bias, td, to)
            exp = golden["q%d_cb%d_endpos" % (i, cbf)]
            assert np.array_equal(got, exp), (i, cbf)
            words += int(exp[:, 3].sum())
    assert words > 50  # the fixture exercises the byte->word promotion (T5)


def test_sw_align_start_positions(golden, oracle):
    td, to = golden["tdata"], golden["toff"]
    for i, q in enumerate(_queries(golden)):
        for cbf in (0, 1):
            cb, bias = oracle.query_cb(q, bool(cbf))
            got = oracle.sw_align(q, cb, bias, td, to)
            assert np.array_equal(got, golden["q%d_cb%d_align" % (i, cbf)]), (i, cbf)


def test_diag_scores(golden, oracle):
    td, to = golden["tdata"], golden["toff"]
    clamped = 0
    for i, q in enumerate(_queries(golden)):
        ids, dg = golden["q%d_diag_ids" % i], golden["q%d_diag_dg" % i]
        for cbf in (0, 1):
            cb4 = oracle.round_bias_diag(golden["q%d_compbias" % i]) if cbf else np.zeros(len(q), np.int8)
            c, r = oracle.diag(q, cb4, td, to, ids, dg)
            assert np.array_equal(c, golden["q%d_cb%d_diag_counts" % (i, cbf)]), (i, cbf)
            assert np.array_equal(r, golden["q%d_cb%d_diag_raw" % (i, cbf)]), (i, cbf)
            clamped += int((c == 255).sum())
    assert clamped > 0  # 255 clamp exercised (T2)


def test_backtrace_and_identities(golden, oracle):
    """A6: banded_sw + computerBacktrace (alignment mode 2 of the reference) for every aligned pair of the fixture"""
    td, to = golden["tdata"], golden["toff"]
    n_gapped = 0
    for i, q in enumerate(_queries(golden)):
        cb, bias = oracle.query_cb(q, True)
        aln = golden["q%d_cb1_align" % i]
        bts, ident = golden["q%d_cb1_bt" % i], golden["q%d_cb1_ident" % i]
        for k in range(len(aln)):
            if aln[k, 4] == -1:
                continue
            t = td[int(to[k]):int(to[k + 1])]
            bt, ids = oracle.backtrace(q, cb, t, aln[k])
            assert bt == str(bts[k]) and ids == ident[k], (i, k)
            n_gapped += ("I" in bt) or ("D" in bt)
    assert n_gapped > 100


def test_fast_scan_restatement_equals_literal(oracle):
    """orc_ungapped_alignment_batch_fast (row-vectorised, used for the at-scale GPU parity test and the "port" CPU baseline) gives the
    scores of the literal restatement orc_ungapped_alignment on random + planted-homolog targets, incl. saturating ones."""
    import os
    from mmseqs2_b200 import synth
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "blosum62.npz"))
    rng = np.random.default_rng(11)
    bg = synth.background(d["pback"])
    res, off = synth.random_seqs(rng, 1500, bg, mean=250, sigma=0.7, lo=1, hi=2500)
    qs = synth.split(*synth.random_seqs(rng, 4, bg, mean=300, sigma=120, lo=1, hi=700, normal=True))
    synth.plant_homologs(rng, res, off, qs, bg, frac=0.1, subst=0.05)
    off64 = off.astype(np.int64)
    saturated = 0
    for q in qs:
        for cbf in (True, False):
            cb, bias = oracle.query_cb(q, cbf)
            a = oracle.ungapped(q, cb, bias, res, off64)
            b = oracle.ungapped(q, cb, bias, res, off64, fast=True)
            assert np.array_equal(a, b)
            saturated += int((a == 255 - bias).sum())
    assert saturated > 0
