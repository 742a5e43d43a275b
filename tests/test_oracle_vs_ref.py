"""CPU, build container only: oracle/oracle.c against the reference's own sources compiled in place
(oracle/_ref/libmmseqs_ref.so).  Skipped where /root/reference is absent (the GPU box)."""
import numpy as np
import pytest

from oracle.pyoracle import Ref, pack_targets
from mmseqs2_b200 import synth

pytestmark = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (no /root/reference here)")


@pytest.fixture(scope="module")
def ref():
    return Ref()


def _targets(rng, bg, q, n):
    tg = []
    for k in range(n):
        m = k % 3
        if m == 0 or len(q) < 4:
            tg.append(synth.random_seqs(rng, 1, bg, mean=200, sigma=0.8, lo=1, hi=900)[0])
        elif m == 1:
            tg.append(synth.mutate(rng, q, bg, 0.37, 0.08))
        else:
            a = int(rng.integers(0, len(q) - 1))
            b = int(rng.integers(a + 1, len(q) + 1))
            pre = synth.random_seqs(rng, 1, bg, mean=20, sigma=0.5, lo=1, hi=60)[0]
            tg.append(np.concatenate([pre, synth.mutate(rng, q[a:b], bg, 0.2, 0.04), pre[::-1]]))
    return pack_targets(tg)


def test_matrix_fixture_matches_reference(ref, blosum):
    mat, pb, _ = ref.matrix()
    assert np.array_equal(mat, blosum[0]) and np.array_equal(pb, blosum[1])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_pairs_all_entry_points(ref, oracle, blosum, seed):
    rng = np.random.default_rng(seed)
    bg = synth.background(blosum[1])
    for it in range(8):
        qL = int(rng.integers(1, 40)) if it % 4 == 0 else int(rng.integers(20, 600))
        q = synth.random_seqs(rng, 1, bg, mean=qL, sigma=0, lo=qL, hi=qL, normal=True)[0]
        if it % 5 == 0:
            q[:: 5] = 20  # X
        td, to = _targets(rng, bg, q, 120)
        f = ref.comp_bias(q)
        assert np.array_equal(f.view(np.uint32), oracle.comp_bias(q).view(np.uint32))
        for cbf in (True, False):
            cb, bias = oracle.query_cb(q, cbf)
            assert np.array_equal(ref.ungapped(q, cbf, td, to), oracle.ungapped(q, cb, bias, td, to))
            assert np.array_equal(ref.sw_score_endpos(q, cbf, td, to), oracle.sw_score_endpos(q, cb, bias, td, to))
            a, _, _ = ref.ssw_align(q, cbf, td, to, mode=1)
            assert np.array_equal(a[:, :6], oracle.sw_align(q, cb, bias, td, to))
        ids = rng.integers(0, len(to) - 1, 2000).astype(np.uint32)
        dg = rng.integers(-qL - 3, 700, 2000).astype(np.int16).view(np.uint16)
        for use in (True, False):
            cb4 = oracle.round_bias_diag(f) if use else np.zeros(qL, np.int8)
            c1, r1 = ref.diag(q, f if use else None, td, to, ids, dg)
            c2, r2 = oracle.diag(q, cb4, td, to, ids, dg)
            assert np.array_equal(c1, c2) and np.array_equal(r1, r2)


def test_gap_penalty_variants(ref, oracle, blosum):
    rng = np.random.default_rng(99)
    bg = synth.background(blosum[1])
    q = synth.random_seqs(rng, 1, bg, mean=180, sigma=0, lo=180, hi=180, normal=True)[0]
    td, to = _targets(rng, bg, q, 90)
    cb, bias = oracle.query_cb(q, True)
    for go, ge in ((11, 1), (10, 2), (5, 5), (20, 1), (2, 1)):
        assert np.array_equal(ref.sw_score_endpos(q, True, td, to, go, ge), oracle.sw_score_endpos(q, cb, bias, td, to, go, ge))
        a, _, _ = ref.ssw_align(q, True, td, to, go, ge, mode=1)
        assert np.array_equal(a[:, :6], oracle.sw_align(q, cb, bias, td, to, go, ge))
