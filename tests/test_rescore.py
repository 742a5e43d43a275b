"""SURVEY 8(f) rows 2 and 3 on the device: b200_diag_score_batch (many queries' hit lists per call) and b200_rescore_diagonal
(DistanceCalculator::computeUngappedAlignment, all five rescore modes) -- against the reference's own outputs (fixtures generated
from oracle/_ref by tests/golden/make_rescore_golden.py) and against the C restatement on seeded random input."""
import os

import numpy as np
import pytest

from mmseqs2_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("score", "start_pos", "end_pos", "diagonal_len", "dist_to_diagonal", "diagonal")


def _ident(q, t, r):
    """rescorediagonal.cpp:296-301: identical letters of the reported segment, case-insensitive"""
    if r["start_pos"] < 0 or r["end_pos"] < r["start_pos"]:
        return 0
    d, dist = int(r["diagonal"]), int(r["dist_to_diagonal"])
    qs, ts = (r["start_pos"] + dist, r["start_pos"]) if d >= 0 else (r["start_pos"], r["start_pos"] + dist)
    n = int(r["end_pos"] - r["start_pos"] + 1)
    return sum((q[qs + k] & 0xDF) == (t[ts + k] & 0xDF) for k in range(n))


def test_rescore_fixture_is_self_consistent(oracle):
    """CPU: the committed reference outputs equal the C restatement (the oracle is pinned on the fixture too)"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "rescore_v1.npz"))
    qo, to = g["qoff"].astype(np.int64), g["toff"].astype(np.int64)
    for mode in range(5):
        for k, (qi, ti, dg) in enumerate(g["hits"]):
            q, t = g["qdata"][qo[qi]:qo[qi + 1]].tobytes(), g["tdata"][to[ti]:to[ti + 1]].tobytes()
            assert np.array_equal(oracle.rescore_diagonal(q, t, int(dg), g["asciimat"], mode), g["expected"][mode][k]), (mode, k)


@pytest.mark.gpu
def test_rescore_diagonal_matches_reference_outputs(ctx):
    g = np.load(os.path.join(ROOT, "tests", "golden", "rescore_v1.npz"))
    qo, to = g["qoff"].astype(np.int64), g["toff"].astype(np.int64)
    queries = [g["qdata"][qo[i]:qo[i + 1]].tobytes() for i in range(len(qo) - 1)]
    ctx.load_db_ascii(g["tdata"], g["toff"])
    hits = g["hits"]
    lists = []
    for qi in range(len(queries)):
        sel = hits[hits[:, 0] == qi]
        lists.append((sel[:, 1].astype(np.uint32), sel[:, 2].astype(np.uint16)))
    for mode in range(5):
        out = ctx.rescore_diagonal(queries, lists, g["asciimat"], mode)
        got = np.stack([out[f] for f in FIELDS], 1).astype(np.int64)
        exp = g["expected"][mode].astype(np.int64).astype(np.int32).astype(np.int64)   # LocalAlignment::score is unsigned: same 32 bits
        assert np.array_equal(got, exp), (mode, np.nonzero((got != exp).any(1))[0][:5])
        for k, (qi, ti, dg) in enumerate(hits):
            exp_id = _ident(queries[qi], g["tdata"][to[ti]:to[ti + 1]].tobytes(), out[k]) if mode >= 2 else 0
            assert out["identical"][k] == exp_id, (mode, k)


@pytest.mark.gpu
def test_rescore_diagonal_random_vs_oracle(ctx, oracle):
    """longer sequences (several 32-cell blocks, block-boundary effects of the warp reductions), long shared segments, every mode"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "rescore_v1.npz"))
    m = g["asciimat"]
    rng = np.random.default_rng(77)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXBZUacdklmn", np.uint8)
    queries, targets, lists = [], [], []
    for qi in range(12):
        qL = int(rng.integers(30, 3000))
        q = bytearray(rng.choice(letters, qL).tobytes())
        queries.append(bytes(q))
        ids, dgs = [], []
        for k in range(25):
            tL = int(rng.integers(1, 3000))
            t = bytearray(rng.choice(letters, tL).tobytes())
            n = int(rng.integers(1, min(qL, tL) + 1)); a = int(rng.integers(0, qL - n + 1)); b = int(rng.integers(0, tL - n + 1))
            t[b:b + n] = q[a:a + n]
            for j in rng.integers(0, n, int(n * rng.uniform(0.0, 0.4))):
                t[b + int(j)] = int(rng.choice(letters))
            if k % 7 == 0:
                t[0] = ord("*")
            ids.append(len(targets)); targets.append(bytes(t))
            dgs.append(((a - b) if k % 3 else int(rng.integers(-tL, qL + 1))) & 0xffff)
        lists.append((np.array(ids, np.uint32), np.array(dgs, np.uint16)))
    toff = np.zeros(len(targets) + 1, np.uint64); toff[1:] = np.cumsum([len(t) for t in targets])
    ctx.load_db_ascii(b"".join(targets), toff)
    for mode in range(5):
        out = ctx.rescore_diagonal(queries, lists, m, mode)
        k = 0
        for qi, (ids, dgs) in enumerate(lists):
            for ti, dg in zip(ids, dgs):
                exp = oracle.rescore_diagonal(queries[qi], targets[int(ti)], int(dg), m, mode).astype(np.int32).astype(np.int64)
                got = np.array([out[f][k] for f in FIELDS], np.int64)
                assert np.array_equal(got, exp), (mode, qi, int(ti), got, exp)
                if mode >= 2:
                    assert out["identical"][k] == _ident(queries[qi], targets[int(ti)], out[k])
                k += 1


@pytest.mark.gpu
def test_diag_score_batch_equals_per_query_calls(ctx, oracle, submat, blosum):
    rng = np.random.default_rng(5)
    bg = synth.background(blosum[1])
    res, off = synth.random_seqs(rng, 4000, bg, mean=250, sigma=0.6, lo=20, hi=3000)
    qs = synth.split(*synth.random_seqs(rng, 9, bg, mean=300, sigma=100, lo=40, hi=700, normal=True))
    synth.plant_homologs(rng, res, off, qs, bg, frac=0.2)
    ctx.load_db(res, off, 21)
    to = off.astype(np.int64)
    profs = [submat.diag_query(q, submat.comp_bias(q)) for q in qs]
    lists = []
    for q in qs:
        n = int(rng.integers(0, 3000))
        lists.append((rng.integers(0, 4000, n).astype(np.uint32), rng.integers(-len(q), 600, n).astype(np.int16).view(np.uint16)))
    got = ctx.diag_score_batch(profs, lists, want_raw=True)
    for qi, q in enumerate(qs):
        c_exp, r_exp = oracle.diag(q, oracle.round_bias_diag(oracle.comp_bias(q)), res, to, lists[qi][0], lists[qi][1])
        assert np.array_equal(got[qi][0], c_exp) and np.array_equal(got[qi][1], r_exp), qi
        if len(lists[qi][0]):
            c1, r1 = ctx.diag_score(profs[qi], lists[qi][0], lists[qi][1], want_raw=True)
            assert np.array_equal(c1, got[qi][0]) and np.array_equal(r1, got[qi][1])
