"""The C++ operator mirror (include/b200_mmseqs.hpp) compiled with g++ -fno-exceptions: syntax on CPU, behaviour on GPU."""
import os
import subprocess

import numpy as np
import pytest

from mmseqs2_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "operators_smoke.cpp")


def test_operator_mirror_compiles_without_exceptions():
    subprocess.check_call(["g++", "-std=c++11", "-fno-exceptions", "-Wall", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), SRC])


@pytest.mark.gpu
def test_operator_mirror_matches_oracle(tmp_path, built_lib, oracle, blosum):
    exe = str(tmp_path / "operators_smoke")
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["g++", "-std=c++11", "-fno-exceptions", "-O1", "-I" + os.path.join(ROOT, "include"), SRC,
                           "-L" + libdir, "-lb200align", "-Wl,-rpath," + libdir, "-o", exe])
    mat, pb = blosum
    rng = np.random.default_rng(31)
    bg = synth.background(pb)
    res, off = synth.random_seqs(rng, 300, bg, mean=200, sigma=0.6, lo=5, hi=900)
    q = synth.random_seqs(rng, 1, bg, mean=280, sigma=0, lo=280, hi=280, normal=True)[0]
    synth.plant_homologs(rng, res, off, [q], bg, frac=0.3, subst=0.25, indel=0.03)
    nh = 800
    ids = rng.integers(0, 300, nh).astype(np.uint32)
    dg = rng.integers(-280, 300, nh).astype(np.int16).view(np.uint16)
    min_score = 40
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        np.array([21, 300, len(q), nh, min_score], np.int32).tofile(f)
        mat.astype(np.int16).tofile(f); pb.astype(np.float64).tofile(f)
        off.astype(np.uint64).tofile(f); res.tofile(f); q.tofile(f); ids.tofile(f); dg.tofile(f)
    subprocess.check_call([exe, inp, outp])
    out = np.fromfile(outp, np.int32)
    full = out[:300 * 8].reshape(300, 8)
    aln = full[:, :6]
    cb, bias = oracle.query_cb(q, True)
    to = off.astype(np.int64)
    exp = oracle.sw_align(q, cb, bias, res, to)
    passed = exp[:, 0] >= min_score
    # hits above the score gate carry full positions; the others only the score (positions never computed)
    assert np.array_equal(aln[:, 0], exp[:, 0])
    assert np.array_equal(aln[passed], exp[passed])
    assert (aln[~passed][:, [1, 3, 4]] == -1).all()
    for k in np.nonzero(passed)[0]:   # alignment mode 2: identities + backtrace length from the device CIGAR
        if exp[k, 4] == -1:
            continue
        bt, n_ident = oracle.backtrace(q, cb, res[int(to[k]):int(to[k + 1])], exp[k])
        assert full[k, 6] == n_ident and full[k, 7] == len(bt), k
    rest = out[300 * 8:]
    n_scan = rest[0]
    scan = rest[1:101].reshape(50, 2)
    dense = oracle.ungapped(q, cb, bias, res, to)
    ids_e = np.nonzero(dense > 15)[0]
    order = np.lexsort((ids_e, -dense[ids_e]))
    ids_e = ids_e[order][:50]
    assert n_scan == len(ids_e)
    assert np.array_equal(scan[:n_scan, 0], ids_e) and np.array_equal(scan[:n_scan, 1], dense[ids_e])
    # GAPLESS_SMITH_WATERMAN: the same 50 targets, rescored with the gapped kernel, ordered (gapped score desc, id asc)
    n_gsw = rest[101]
    gsw = rest[102:302].reshape(50, 4)
    assert n_gsw == n_scan
    from oracle.pyoracle import pack_targets
    sd, so = pack_targets([res[int(to[t]):int(to[t + 1])] for t in ids_e])
    e = oracle.sw_score_endpos(q, cb, bias, sd, so)
    order2 = np.lexsort((ids_e, -e[:, 0]))
    exp_gsw = np.stack([ids_e[order2], e[order2, 0], e[order2, 1], e[order2, 2]], 1)
    assert np.array_equal(gsw[:n_gsw], exp_gsw)
    rest = rest[201:]
    counts = rest[101:101 + nh]
    c_exp, _ = oracle.diag(q, oracle.round_bias_diag(oracle.comp_bias(q)), res, to, ids, dg)
    assert np.array_equal(counts, c_exp.astype(np.int32))


ALIGN_SRC = os.path.join(ROOT, "tests", "cpp", "align_batch_smoke.cpp")


def test_align_batch_host_compiles_without_exceptions():
    subprocess.check_call(["g++", "-std=c++11", "-fno-exceptions", "-Wall", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), ALIGN_SRC])


@pytest.mark.gpu
def test_align_batch_from_cpp_host_writes_the_reference_entries(tmp_path, built_lib, blosum):
    """b200_align_batch + b200h_result_to_buffer called from a C++ host: the entries equal the reference's `align -a` output"""
    exe = str(tmp_path / "align_batch_smoke")
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["g++", "-std=c++11", "-fno-exceptions", "-O1", "-I" + os.path.join(ROOT, "include"), ALIGN_SRC,
                           "-L" + libdir, "-lb200align", "-Wl,-rpath," + libdir, "-o", exe])
    mat, pb = blosum
    ex = np.load(os.path.join(ROOT, "tests", "golden", "examples_v1.npz"))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "align_v1.npz"))
    nq, nt = len(ex["qoff"]) - 1, len(ex["toff"]) - 1
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        np.array([21, nt, nq, len(gold["hit_targets"]), 2], np.int32).tofile(f)
        mat.astype(np.int16).tofile(f); pb.astype(np.float64).tofile(f)
        ex["toff"].astype(np.uint64).tofile(f); ex["tdata"].tofile(f)
        ex["qoff"].astype(np.uint64).tofile(f); ex["qdata"].tofile(f)
        gold["hit_off"].astype(np.uint64).tofile(f); gold["hit_targets"].astype(np.uint32).tofile(f)
        gold["target_keys"].astype(np.uint32).tofile(f); (5000 + np.arange(nq)).astype(np.uint32).tofile(f)
    subprocess.check_call([exe, inp, outp])
    entries = open(outp, "rb").read().split(b"\0")[:nq]
    for qi in range(nq):
        assert entries[qi] == bytes(gold["cfg_default_a_text"][qi]), qi
