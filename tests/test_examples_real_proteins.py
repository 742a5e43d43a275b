"""BASELINE config[0] data in miniature: real proteins from the reference's examples/ (first 40 queries x first 600 targets,
numeric residues as the reference maps them, incl. X/B/Z/U letters and low-complexity regions).  The committed outputs are
the reference's own ungapped_alignment and ssw_align(mode 1) results.  CPU: oracle; GPU: the C ABI."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ex():
    return np.load(os.path.join(ROOT, "tests", "golden", "examples_v1.npz"))


def _seqs(d, o):
    return [d[int(o[i]):int(o[i + 1])] for i in range(len(o) - 1)]


def test_oracle_on_real_proteins(ex, oracle):
    qs = _seqs(ex["qdata"], ex["qoff"])
    td, to = ex["tdata"], ex["toff"]
    for qi in range(0, len(qs), 4):
        cb, bias = oracle.query_cb(qs[qi], True)
        assert np.array_equal(oracle.ungapped(qs[qi], cb, bias, td, to), ex["ungapped"][qi].astype(np.int32)), qi
        assert np.array_equal(oracle.sw_align(qs[qi], cb, bias, td, to), ex["align"][qi]), qi


@pytest.mark.gpu
def test_gpu_on_real_proteins(ex, ctx, submat):
    qs = _seqs(ex["qdata"], ex["qoff"])
    n = len(ex["toff"]) - 1
    ctx.load_db(ex["tdata"], ex["toff"].astype(np.uint64), 21)
    profs = [submat.ssw_query(q) for q in qs]
    _, _, dense = ctx.ungapped_scan(profs, want_dense=True)
    assert np.array_equal(dense, ex["ungapped"])
    pairs = np.array([(qi, t) for qi in range(len(qs)) for t in range(n)], np.uint32)
    aln = ctx.sw_align(profs, pairs)
    got = np.stack([aln[f] for f in ("score", "qstart", "qend", "dbstart", "dbend", "word")], 1).reshape(len(qs), n, 6)
    bad = np.nonzero((got != ex["align"]).any(2))
    assert len(bad[0]) == 0, (bad[0][:5], bad[1][:5])
    assert np.array_equal(ctx.sw_score(profs, pairs).reshape(len(qs), n), ex["align"][:, :, 0])


@pytest.mark.gpu
def test_search_hit_lists_end_to_end(ex, ctx, submat):
    """prefilter (ungapped scan, >15, top 300) -> gapped score -> E-value gate (as a per-query minimum raw score, taken from
    the reference's EvalueComputation at fixture time) -> positions -> Matcher::compareHits order == the reference's hit lists"""
    qs = _seqs(ex["qdata"], ex["qoff"])
    tlen = np.diff(ex["toff"])
    ctx.load_db(ex["tdata"], ex["toff"].astype(np.uint64), 21)
    profs = [submat.ssw_query(q) for q in qs]
    hits, n_hits, _ = ctx.ungapped_scan(profs, min_score_excl=15, max_hits=300)
    pairs = np.array([(qi, int(t)) for qi in range(len(qs)) for t in hits[qi]["id"][:int(n_hits[qi])]], np.uint32)
    score = ctx.sw_score(profs, pairs)
    gate = (score >= ex["min_score"][pairs[:, 0]]).astype(np.uint8)
    aln = ctx.sw_align(profs, pairs, gate=gate)
    total = 0
    for qi in range(len(qs)):
        sel = np.nonzero((pairs[:, 0] == qi) & (gate == 1) & (aln["dbend"] != -1))[0]
        order = sorted(sel, key=lambda k: (-int(aln["score"][k]), int(tlen[pairs[k, 1]]), int(pairs[k, 1])))
        exp_ids = ex["hit_ids"][int(ex["hit_off"][qi]):int(ex["hit_off"][qi + 1])]
        exp_rows = ex["hit_rows"][int(ex["hit_off"][qi]):int(ex["hit_off"][qi + 1])]
        assert [int(pairs[k, 1]) for k in order] == exp_ids.tolist(), qi
        got = np.array([[aln[f][k] for f in ("score", "qstart", "qend", "dbstart", "dbend", "word")] for k in order], np.int32).reshape(-1, 6)
        assert np.array_equal(got, exp_rows), qi
        total += len(order)
    assert total == len(ex["hit_ids"]) and total > 100
