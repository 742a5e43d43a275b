"""`rescorediagonal` over DB files (include/b200_db.h: b200_rescorediagonal_db) against the DBs the reference binary writes
(tests/golden/make_rescore_module_golden.py: eight parameter sets, result DB walked in data order, one run of a DB against itself).
On CPU the module's host half runs with the oracle as the per-hit scorer (b200h_rescorediagonal_db_with); on a GPU the product entry
point runs with the device scorer.  Both must reproduce the reference's files byte for byte."""
import ctypes
import importlib.util
import os

import numpy as np
import pytest

from mmseqs2_b200 import db

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUFFIXES = ("", ".index", ".dbtype")


def _param_sets():
    spec = importlib.util.spec_from_file_location("make_rescore_module_golden", os.path.join(ROOT, "tests", "golden", "make_rescore_module_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.PARAM_SETS


PARAM_SETS = _param_sets()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "rescore_module_v1.npz"), allow_pickle=False)


@pytest.fixture(scope="module")
def alphabet():
    return np.load(os.path.join(ROOT, "tests", "golden", "blosum62.npz"))["alphabet"].tobytes()


@pytest.fixture()
def dbs(gold, tmp_path):
    for name in ("Q", "T", "pref", "pref_tt"):
        for suf in SUFFIXES:
            open(str(tmp_path / name) + suf, "wb").write(gold[name + suf].tobytes())
    return tmp_path


def oracle_scorer(orc, asciimat):
    """a b200_rescore_fn that answers with the oracle's computeUngappedAlignment and the identity count of rescorediagonal.cpp:296-301"""
    state = {}

    def arr(ptr, ctype, n):
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), (n,))

    def fn(user, qd, qo, nq, ho, ids, dg, td, to, nt, m, mode, out):
        try:
            if mode < 0:
                state["to"] = arr(to, ctypes.c_uint64, nt + 1).copy()
                state["td"] = ctypes.string_at(td, int(state["to"][-1]))
                return 0
            qo_, ho_ = arr(qo, ctypes.c_uint64, nq + 1), arr(ho, ctypes.c_uint64, nq + 1)
            n = int(ho_[nq])
            ids_, dg_ = arr(ids, ctypes.c_uint32, n), arr(dg, ctypes.c_uint16, n)
            out_ = arr(out, ctypes.c_int32, n * 7).reshape(n, 7)
            qd_ = ctypes.string_at(qd, int(qo_[nq]))
            for qi in range(nq):
                q = qd_[int(qo_[qi]):int(qo_[qi + 1])]
                for h in range(int(ho_[qi]), int(ho_[qi + 1])):
                    t = state["td"][int(state["to"][ids_[h]]):int(state["to"][ids_[h] + 1])]
                    out_[h, :6] = np.asarray(orc.rescore_diagonal(q, t, int(dg_[h]), asciimat, mode), np.int64).astype(np.int32)
                    ident = 0
                    if mode >= 2:
                        st, en, dist, diag = int(out_[h, 1]), int(out_[h, 2]), int(out_[h, 4]), int(out_[h, 5])
                        qs, ds = (st + dist, st) if diag >= 0 else (st, st + dist)
                        for k in range(en - st + 1):
                            if qs + k < len(q) and ds + k < len(t) and (q[qs + k] & 0xdf) == (t[ds + k] & 0xdf):
                                ident += 1
                    out_[h, 6] = ident
            return 0
        except Exception:   # pragma: no cover -- an exception must not unwind through the C caller
            import traceback
            traceback.print_exc()
            return 3
    return db.RESCORE_FN(fn)


def _check(gold, tmp_path, name):
    for suf in SUFFIXES:
        assert open(str(tmp_path / ("M_" + name)) + suf, "rb").read() == gold["out_" + name + suf].tobytes(), (name, suf)


def test_ascii_matrix_equals_the_reference_table(submat, alphabet):
    g = np.load(os.path.join(ROOT, "tests", "golden", "rescore_v1.npz"))
    assert np.array_equal(db.ascii_matrix(submat, alphabet), g["asciimat"].reshape(123, 123))


@pytest.mark.parametrize("name", sorted(PARAM_SETS))
def test_host_half_with_the_oracle_scorer_reproduces_the_reference_db(gold, dbs, oracle, submat, alphabet, name):
    flags, kw, same = PARAM_SETS[name]
    q, p = ("T", "pref_tt") if same else ("Q", "pref")
    scorer = oracle_scorer(oracle, db.ascii_matrix(submat, alphabet))
    nh, nr = db.rescorediagonal_db(None, submat, alphabet, str(dbs / q), str(dbs / "T"), str(dbs / p), str(dbs / ("M_" + name)),
                                   db.RescoreParams(**kw), scorer=scorer, bucket_queries=7)
    assert nh > 0
    _check(gold, dbs, name)


def test_module_refuses_what_it_does_not_cover(dbs, oracle, submat, alphabet):
    scorer = oracle_scorer(oracle, db.ascii_matrix(submat, alphabet))
    with pytest.raises(db.B200Error):
        db.rescorediagonal_db(None, submat, alphabet, str(dbs / "Q"), str(dbs / "T"), str(dbs / "pref"), str(dbs / "o"), db.RescoreParams(rescore_mode=7), scorer=scorer)
    with pytest.raises(db.B200Error):
        db.rescorediagonal_db(None, submat, alphabet, str(dbs / "missing"), str(dbs / "T"), str(dbs / "pref"), str(dbs / "o"), db.RescoreParams(), scorer=scorer)
    open(str(dbs / "T.dbtype"), "wb").write(np.array([1], np.int32).tobytes())          # a nucleotide DB
    with pytest.raises(db.B200Error):
        db.rescorediagonal_db(None, submat, alphabet, str(dbs / "Q"), str(dbs / "T"), str(dbs / "pref"), str(dbs / "o"), db.RescoreParams(), scorer=scorer)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(PARAM_SETS))
def test_module_on_the_device_reproduces_the_reference_db(gold, dbs, ctx, submat, alphabet, name):
    flags, kw, same = PARAM_SETS[name]
    q, p = ("T", "pref_tt") if same else ("Q", "pref")
    nh, nr = db.rescorediagonal_db(ctx, submat, alphabet, str(dbs / q), str(dbs / "T"), str(dbs / p), str(dbs / ("M_" + name)),
                                   db.RescoreParams(**kw), bucket_queries=7)
    assert nh > 0
    _check(gold, dbs, name)
