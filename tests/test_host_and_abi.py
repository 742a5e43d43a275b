"""CPU: host-side profile logic of the product (include/b200_host.h) against the oracle, and the C-ABI surface:
libb200align.so loads and exports every symbol the headers declare (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200h?_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared("b200_align.h") + _declared("b200_host.h") + _declared("b200_alignment.h") + _declared("b200_db.h") + _declared("b200_gpuserver.h") + _declared("b200_multi.h")
    assert len(names) >= 53
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from mmseqs2_b200 import api
    monkeypatch.setattr(api, "_LIB", None)
    monkeypatch.setattr(api, "_HERE", str(tmp_path))
    try:
        api.load_library()
    except api.B200Error as e:
        assert "no CPU path" in str(e)
    else:
        raise AssertionError("expected B200Error")


def test_host_profiles_match_oracle(golden, oracle, submat):
    for i in range(int(golden["nq"])):
        q = golden["q%d" % i]
        f = submat.comp_bias(q)
        assert np.array_equal(f.view(np.uint32), golden["q%d_compbias" % i].view(np.uint32))
        for cbf in (True, False):
            qp = submat.ssw_query(q, comp_bias=cbf)
            cb, bias = oracle.query_cb(q, cbf)
            assert qp.bias == bias and np.array_equal(qp.cb, cb)
            exp = oracle.mat[:, q].astype(np.int32) + cb.astype(np.int32)[None, :]   # mat[a][q[j]] + cb[j]
            assert np.array_equal(qp.profile.astype(np.int32), exp)
        dq = submat.diag_query(q, f)
        cb4 = oracle.round_bias_diag(f)
        exp = oracle.mat[q, :].T.astype(np.int32) + cb4.astype(np.int32)[None, :]     # mat[q[j]][a] + cb4[j]
        assert np.array_equal(dq.profile.astype(np.int32), exp)


def test_bias_recovered_from_profile(golden, oracle, submat, built_lib):
    """b200h_ssw_bias_from_profile (what the Marv shim and the gpuserver use): a sequence query's profile gives back ssw_init's
    bias; a table that is not matrix column + one bias per position falls under the profile-query rule."""
    lib = ctypes.CDLL(built_lib)
    lib.b200h_ssw_bias_from_profile.restype = ctypes.c_int
    mat16 = np.ascontiguousarray(oracle.mat.astype(np.int16))
    A = mat16.shape[0]
    for i in range(int(golden["nq"])):
        q = np.ascontiguousarray(golden["q%d" % i])
        for cbf in (True, False):
            qp = submat.ssw_query(q, comp_bias=cbf)
            prof = np.ascontiguousarray(qp.profile)
            got = lib.b200h_ssw_bias_from_profile(mat16.ctypes.data_as(ctypes.c_void_p), A, q.ctypes.data_as(ctypes.c_void_p), len(q),
                                                  prof.ctypes.data_as(ctypes.c_void_p))
            assert got == qp.bias
    rng = np.random.default_rng(5)
    pssm = rng.integers(-9, 12, (A, 77)).astype(np.int8)
    pssm[A - 1] = 0
    q = rng.integers(0, 20, 77).astype(np.uint8)
    got = lib.b200h_ssw_bias_from_profile(mat16.ctypes.data_as(ctypes.c_void_p), A, q.ctypes.data_as(ctypes.c_void_p), 77,
                                          pssm.ctypes.data_as(ctypes.c_void_p))
    assert got == -int(pssm[:A - 1].min())


def test_marv_shim_compiles_without_exceptions():
    """integration/shim/marv.h is what the patched reference host compiles against (-fno-exceptions, C++14)."""
    import subprocess
    src = '#include "marv.h"\nint main() { return sizeof(Marv::Result) == 16 ? 0 : 1; }\n'
    cmd = ["g++", "-std=c++14", "-fno-exceptions", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "integration", "shim"),
           "-I" + os.path.join(ROOT, "include"), "-x", "c++", "-"]
    r = subprocess.run(cmd, input=src.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
