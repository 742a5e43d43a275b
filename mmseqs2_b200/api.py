"""ctypes binding of libb200align.so, shaped after the reference operator surface.

Reference interface            ->  here
  SubstitutionMatrix / BaseMatrix (subMatrix, pBack)                 SubMatrix
  SmithWaterman::ssw_init (StripedSmithWaterman.cpp:1364)            SubMatrix.ssw_query()      -> QueryProfile
  UngappedAlignment::createProfile (UngappedAlignment.cpp:388)       SubMatrix.diag_query()     -> QueryProfile
  Marv::loadDb/setDb (marv.h:20-24), SequenceLookup                  Context.load_db()
  Marv::scan / SmithWaterman::ungapped_alignment over the DB         Context.ungapped_scan()
  UngappedAlignment::align / scoreSingleSequence                     Context.diag_score()
  SmithWaterman::alignScoreEndPos                                    Context.sw_score_endpos()
  SmithWaterman::ssw_align (modes 0/1)                               Context.sw_align()
  banded_sw + computerBacktrace (alignment mode 3)                   Context.sw_backtrace()
  BandedNucleotideAligner::align (ksw_extz2)                         Context.nucl_align()
  DistanceCalculator::computeUngappedAlignment (rescorediagonal)     Context.load_db_ascii() + Context.rescore_diagonal()
  Marv over several GPUs (cudasw4 partitions + merge)                MultiContext
  Alignment::run for a bucket of queries                             mmseqs2_b200.alignment.align_batch()
  the modules over DB files                                          mmseqs2_b200.db
"""
import ctypes
import weakref
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_vp = ctypes.c_void_p
_u64 = ctypes.c_uint64


class B200Error(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libb200align.so")


def load_library():
    """Load libb200align.so; fail loudly if it was not built (no fallback of any kind)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise B200Error("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the CUDA library is the product; there is no CPU path)" % p)
    lib = ctypes.CDLL(p)
    lib.b200_last_error.restype = ctypes.c_char_p
    lib.b200_last_error.argtypes = [_vp]
    for name in ("b200_launch_count", "b200_db_num_seqs", "b200_db_num_residues", "b200_job_cells"):
        getattr(lib, name).restype = _u64
        getattr(lib, name).argtypes = [_vp]
    lib.b200_destroy.argtypes = [_vp]
    lib.b200_destroy.restype = None
    lib.b200_job_destroy.argtypes = [_vp]
    lib.b200_job_destroy.restype = None
    _LIB = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


class _CQuery(ctypes.Structure):
    _fields_ = [("profile", _vp), ("qlen", ctypes.c_int32), ("bias", ctypes.c_int32)]


HIT_DTYPE = np.dtype([("id", np.uint32), ("score", np.int32)])
PAIR_DTYPE = np.dtype([("query", np.uint32), ("target", np.uint32)])
END_DTYPE = np.dtype([("score", np.int32), ("qend", np.int32), ("dbend", np.int32), ("word", np.int32)])
RESCORE_DTYPE = np.dtype([("score", np.int32), ("start_pos", np.int32), ("end_pos", np.int32), ("diagonal_len", np.int32),
                          ("dist_to_diagonal", np.int32), ("diagonal", np.int32), ("identical", np.int32)])
NUCL_TASK_DTYPE = np.dtype([("query", np.uint32), ("target", np.uint32), ("diagonal", np.uint16), ("reserved", np.uint16)])
NUCL_ALN_DTYPE = np.dtype([("score", np.int32), ("qstart", np.int32), ("qend", np.int32), ("dbstart", np.int32), ("dbend", np.int32),
                           ("identical", np.int32), ("n_cigar", np.int32)])
ALN_DTYPE = np.dtype([("score", np.int32), ("qstart", np.int32), ("qend", np.int32), ("dbstart", np.int32),
                      ("dbend", np.int32), ("word", np.int32)])


class QueryProfile:
    """int8 [A][qlen] profile + the SSW bias constant (what ssw_init leaves in s_profile)."""

    def __init__(self, profile, bias=0, cb=None):
        self.profile = np.ascontiguousarray(profile, np.int8)
        self.A, self.qlen = self.profile.shape
        self.bias = int(bias)
        self.cb = cb


class SubMatrix:
    """Integer substitution matrix + background; builds query profiles with the reference's rounding rules (host)."""

    def __init__(self, mat, pback):
        self.lib = load_library()
        self.mat = np.ascontiguousarray(mat, np.int16)
        self.pback = np.ascontiguousarray(pback, np.float64)
        self.A = int(self.mat.shape[0])
        self.lib.b200h_ssw_bias.restype = ctypes.c_int
        self.lib.b200h_build_profile.restype = ctypes.c_int

    def comp_bias(self, q, scale=1.0):
        q = np.ascontiguousarray(q, np.uint8)
        out = np.zeros(len(q), np.float32)
        self.lib.b200h_comp_bias(_p(self.mat), _p(self.pback), self.A, _p(q), len(q), ctypes.c_float(scale), _p(out))
        return out

    def ssw_query(self, q, comp_bias=True, scale=1.0):
        """SmithWaterman::ssw_init for a sequence query: profile[a][j] = mat[a][q[j]] + cb[j], bias constant."""
        q = np.ascontiguousarray(q, np.uint8)
        cb = np.zeros(len(q), np.int8)
        if comp_bias:
            f = self.comp_bias(q, scale)
            self.lib.b200h_round_bias_ssw(_p(f), len(q), _p(cb))
        bias = self.lib.b200h_ssw_bias(_p(self.mat), self.A, _p(cb), len(q), 1 if comp_bias else 0)
        prof = np.zeros((self.A, len(q)), np.int8)
        if self.lib.b200h_build_profile(_p(self.mat), self.A, _p(q), len(q), _p(cb), 1, _p(prof)) != 0:
            raise B200Error("profile value outside int8")
        return QueryProfile(prof, bias, cb)

    def pssm_query(self, pssm):
        """Profile (PSSM) query, the HMM_PROFILE branch of ssw_init / createProfile: pssm = int8 [20][L] as
        Sequence::getAlignmentProfile() holds it.  The same [A][L] table serves the scan, the gapped kernels and the
        per-diagonal scorer."""
        pssm = np.ascontiguousarray(pssm, np.int8)
        rows, L = pssm.shape
        prof = np.zeros((self.A, L), np.int8)
        bias = self.lib.b200h_build_profile_pssm(_p(pssm), rows, L, self.A, _p(prof))
        if bias < 0:
            raise B200Error("bad PSSM shape")
        return QueryProfile(prof, bias, np.zeros(L, np.int8))

    def diag_query(self, q, bias_f32=None):
        """UngappedAlignment::createProfile: profile[a][j] = mat[q[j]][a] + round(bias[j]/4)."""
        q = np.ascontiguousarray(q, np.uint8)
        cb = np.zeros(len(q), np.int8)
        if bias_f32 is not None:
            f = np.ascontiguousarray(bias_f32, np.float32)
            self.lib.b200h_round_bias_diag(_p(f), len(q), _p(cb))
        prof = np.zeros((self.A, len(q)), np.int8)
        if self.lib.b200h_build_profile(_p(self.mat), self.A, _p(q), len(q), _p(cb), 0, _p(prof)) != 0:
            raise B200Error("profile value outside int8")
        return QueryProfile(prof, 0, cb)


def _cqueries(queries):
    arr = (_CQuery * len(queries))()
    for i, q in enumerate(queries):
        arr[i].profile = q.profile.ctypes.data
        arr[i].qlen = q.qlen
        arr[i].bias = q.bias
    return arr


class Job:
    def __init__(self, ctx, handle, kind, nq=0, k=0, n=0):
        self.ctx, self.handle, self.kind, self.nq, self.k, self.n = ctx, handle, kind, nq, k, n
        ctx._jobs.add(self)        # a job must not outlive its context: Context.close() destroys what is still open

    def run(self):
        self.ctx._check(self.ctx.lib.b200_job_run(self.handle))

    @property
    def cells(self):
        return int(self.ctx.lib.b200_job_cells(self.handle))

    def fetch(self, want_dense=False):
        ctx = self.ctx
        if self.kind == "scan":
            hits = np.zeros((self.nq, self.k), HIT_DTYPE)
            n_hits = np.zeros(self.nq, np.uint32)
            dense = np.zeros((self.nq, ctx.n_seq), np.uint8) if want_dense else None
            ctx._check(ctx.lib.b200_scan_job_fetch(self.handle, _p(hits), _p(n_hits), _p(dense)))
            return hits, n_hits, dense
        if self.kind == "sw_score":
            out = np.zeros(self.n, np.int32)
            ctx._check(ctx.lib.b200_sw_score_job_fetch(self.handle, _p(out)))
            return out
        out = np.zeros(self.n, END_DTYPE)
        ctx._check(ctx.lib.b200_sw_job_fetch(self.handle, _p(out)))
        return out

    def close(self):
        if self.handle:
            if getattr(self.ctx, "h", None):      # the context is still alive (b200_job_destroy touches its stream and mutex)
                self.ctx.lib.b200_job_destroy(self.handle)
            self.handle = None
            self.ctx._jobs.discard(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One GPU, one stream, one resident target DB (the role of class Marv / SequenceLookup)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.b200_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise B200Error("b200_create(%d) failed with %d (is a CUDA device visible?)" % (device, rc))
        self.h = h
        self.n_seq = 0
        self._jobs = weakref.WeakSet()

    def close(self):
        if getattr(self, "h", None):
            for j in list(self._jobs):
                j.close()
            self.lib.b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise B200Error("b200 error %d: %s" % (rc, self.lib.b200_last_error(self.h).decode()))

    def device_info(self):
        sm, ma, mi, hbm = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), _u64()
        self._check(self.lib.b200_device_info(self.h, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(hbm)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "hbm_bytes": hbm.value}

    @property
    def launches(self):
        return int(self.lib.b200_launch_count(self.h))

    @property
    def last_kernel_ms(self):
        self.lib.b200_last_kernel_ms.restype = ctypes.c_float
        self.lib.b200_last_kernel_ms.argtypes = [_vp]
        return float(self.lib.b200_last_kernel_ms(self.h))

    def sync(self):
        self._check(self.lib.b200_sync(self.h))

    def event_record(self, slot):
        self._check(self.lib.b200_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = ctypes.c_float()
        self._check(self.lib.b200_event_elapsed_ms(self.h, a, b, ctypes.byref(ms)))
        return ms.value

    # ---- DB
    def load_db(self, residues, offsets, alphabet):
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        self._check(self.lib.b200_db_load(self.h, _p(residues), _p(offsets), _u64(len(offsets) - 1), int(alphabet)))
        self.n_seq = len(offsets) - 1
        self._db_len = np.diff(offsets.astype(np.int64))

    def db_lengths(self):
        """lengths of the loaded target sequences (host copy kept at load time)"""
        return self._db_len

    # ---- A2
    def ungapped_scan(self, queries, min_score_excl=15, max_hits=300, want_dense=False):
        cq = _cqueries(queries)
        nq = len(queries)
        hits = np.zeros((nq, max_hits), HIT_DTYPE)
        n_hits = np.zeros(nq, np.uint32)
        dense = np.zeros((nq, self.n_seq), np.uint8) if want_dense else None
        self._check(self.lib.b200_ungapped_scan(self.h, cq, nq, int(min_score_excl), ctypes.c_uint32(max_hits), _p(hits),
                                                _p(n_hits), _p(dense)))
        return hits, n_hits, dense

    def scan_job(self, queries, min_score_excl=15, max_hits=300):
        cq = _cqueries(queries)
        h = _vp()
        self._check(self.lib.b200_scan_job_create(self.h, cq, len(queries), int(min_score_excl), ctypes.c_uint32(max_hits),
                                                  ctypes.byref(h)))
        return Job(self, h, "scan", nq=len(queries), k=max_hits)

    # ---- A1
    def diag_score(self, query, ids, diagonals, counts=None, want_raw=False):
        ids = np.ascontiguousarray(ids, np.uint32)
        dg = np.ascontiguousarray(diagonals, np.uint16)
        cnt = np.zeros(len(ids), np.uint8) if counts is None else np.ascontiguousarray(counts, np.uint8).copy()
        raw = np.zeros(len(ids), np.int32) if want_raw else None
        cq = _cqueries([query])
        self._check(self.lib.b200_diag_score(self.h, cq, _p(ids), _p(dg), _u64(len(ids)), _p(cnt), _p(raw)))
        return cnt, raw

    def diag_score_batch(self, queries, hit_lists, want_raw=False):
        """queries[i] with hit_lists[i] = (ids, diagonals): one call, one launch (b200_diag_score_batch) -> [(counts, raw)]"""
        off = np.zeros(len(queries) + 1, np.uint64)
        off[1:] = np.cumsum([len(h[0]) for h in hit_lists])
        ids = np.ascontiguousarray(np.concatenate([np.asarray(h[0], np.uint32) for h in hit_lists]), np.uint32)
        dg = np.ascontiguousarray(np.concatenate([np.asarray(h[1], np.uint16) for h in hit_lists]), np.uint16)
        cnt = np.zeros(len(ids), np.uint8)
        raw = np.zeros(len(ids), np.int32) if want_raw else None
        cq = _cqueries(queries)
        self._check(self.lib.b200_diag_score_batch(self.h, cq, len(queries), _p(off), _p(ids), _p(dg), _p(cnt), _p(raw)))
        o = off.astype(np.int64)
        return [(cnt[o[i]:o[i + 1]], raw[o[i]:o[i + 1]] if want_raw else None) for i in range(len(queries))]

    # ---- rescorediagonal
    def load_db_ascii(self, data, offsets):
        data = np.ascontiguousarray(np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else data, np.uint8)
        off = np.ascontiguousarray(offsets, np.uint64)
        self._check(self.lib.b200_db_load_ascii(self.h, _p(data), _p(off), _u64(len(off) - 1)))

    def rescore_diagonal(self, queries_ascii, hit_lists, asciimat, mode):
        """queries_ascii[i] (bytes) with hit_lists[i] = (ids, diagonals u16) -> structured array per hit (b200_rescore)"""
        qd = np.frombuffer(b"".join(queries_ascii), np.uint8)
        qo = np.zeros(len(queries_ascii) + 1, np.uint64)
        qo[1:] = np.cumsum([len(q) for q in queries_ascii])
        ho = np.zeros(len(queries_ascii) + 1, np.uint64)
        ho[1:] = np.cumsum([len(h[0]) for h in hit_lists])
        ids = np.ascontiguousarray(np.concatenate([np.asarray(h[0], np.uint32) for h in hit_lists]), np.uint32)
        dg = np.ascontiguousarray(np.concatenate([np.asarray(h[1], np.uint16) for h in hit_lists]), np.uint16)
        out = np.zeros(len(ids), RESCORE_DTYPE)
        m = np.ascontiguousarray(asciimat, np.int8)
        self._check(self.lib.b200_rescore_diagonal(self.h, _p(np.ascontiguousarray(qd)), _p(qo), len(queries_ascii), _p(ho), _p(ids), _p(dg),
                                                   _p(m), int(mode), _p(out)))
        return out

    # ---- A3-A5
    @staticmethod
    def _pairs(pairs):
        arr = np.zeros(len(pairs), PAIR_DTYPE)
        pairs = np.asarray(pairs)
        if len(pairs):
            arr["query"] = pairs[:, 0]
            arr["target"] = pairs[:, 1]
        return arr

    def sw_score_endpos(self, queries, pairs, go=11, ge=1):
        cq = _cqueries(queries)
        pa = pairs if getattr(pairs, "dtype", None) == PAIR_DTYPE else self._pairs(pairs)
        out = np.zeros(len(pa), END_DTYPE)
        self._check(self.lib.b200_sw_score_endpos(self.h, cq, len(queries), _p(pa), _u64(len(pa)), go, ge, _p(out)))
        return out

    def sw_score(self, queries, pairs, go=11, ge=1):
        cq = _cqueries(queries)
        pa = pairs if getattr(pairs, "dtype", None) == PAIR_DTYPE else self._pairs(pairs)
        out = np.zeros(len(pa), np.int32)
        self._check(self.lib.b200_sw_score(self.h, cq, len(queries), _p(pa), _u64(len(pa)), go, ge, _p(out)))
        return out

    def sw_score_job(self, queries, pairs, go=11, ge=1):
        cq = _cqueries(queries)
        pa = pairs if getattr(pairs, "dtype", None) == PAIR_DTYPE else self._pairs(pairs)
        h = _vp()
        self._check(self.lib.b200_sw_score_job_create(self.h, cq, len(queries), _p(pa), _u64(len(pa)), go, ge, ctypes.byref(h)))
        return Job(self, h, "sw_score", n=len(pa))

    def sw_startpos(self, queries, pairs, ends, go=11, ge=1):
        cq = _cqueries(queries)
        pa = pairs if getattr(pairs, "dtype", None) == PAIR_DTYPE else self._pairs(pairs)
        ends = np.ascontiguousarray(ends, END_DTYPE)
        out = np.zeros(len(pa), ALN_DTYPE)
        self._check(self.lib.b200_sw_startpos(self.h, cq, len(queries), _p(pa), _u64(len(pa)), go, ge, _p(ends), _p(out)))
        return out

    def sw_align(self, queries, pairs, go=11, ge=1, gate=None):
        cq = _cqueries(queries)
        pa = pairs if getattr(pairs, "dtype", None) == PAIR_DTYPE else self._pairs(pairs)
        g = None if gate is None else np.ascontiguousarray(gate, np.uint8)
        out = np.zeros(len(pa), ALN_DTYPE)
        self._check(self.lib.b200_sw_align(self.h, cq, len(queries), _p(pa), _u64(len(pa)), go, ge, _p(g), _p(out)))
        return out

    # ---- A6
    def sw_backtrace(self, queries, query_seqs, pairs, alns, go=11, ge=1):
        """-> (structured results, list of backtrace strings); alns as returned by sw_align"""
        cq = _cqueries(queries)
        pa = pairs if getattr(pairs, "dtype", None) == PAIR_DTYPE else self._pairs(pairs)
        alns = np.ascontiguousarray(alns, ALN_DTYPE)
        seqs = [np.ascontiguousarray(x, np.uint8) for x in query_seqs]
        ptrs = (_vp * len(seqs))(*[x.ctypes.data for x in seqs])
        qlen = np.array([len(x) for x in seqs], np.int64)
        tl = np.maximum(alns["dbend"] - alns["dbstart"] + 1, 0).astype(np.int64)
        ql = np.maximum(alns["qend"] - alns["qstart"] + 1, 0).astype(np.int64)
        slots = (ql + tl + 2).astype(np.uint64)
        coff = np.zeros(len(pa) + 1, np.uint64)
        coff[1:] = np.cumsum(slots)
        cig = np.zeros(int(coff[-1]) + 1, np.uint32)
        out = np.zeros(len(pa), np.dtype([("n_cigar", np.int32), ("identical", np.int32), ("bt_len", np.int32), ("ok", np.int32)]))
        self._check(self.lib.b200_sw_backtrace(self.h, cq, ptrs, len(queries), _p(pa), _u64(len(pa)), go, ge, _p(alns), _p(out), _p(cig),
                                               _p(coff)))
        bts = []
        for i in range(len(pa)):
            cg = cig[int(coff[i]):int(coff[i]) + int(out["n_cigar"][i])]
            bts.append("".join("MID"[int(c & 0xf)] * int(c >> 4) for c in cg))
        return out, bts

    # ---- A7
    def nucl_align(self, queries, tasks, go=5, ge=2, zdrop=40, decode=True):
        """queries: list of uint8 arrays (A,C,T,G,X = 0..4); tasks: iterable of (query, target, diagonal_u16).
        -> (structured results, list of cigar op arrays, list of backtrace strings)"""
        if isinstance(queries, tuple):          # already packed: (residues uint8, offsets uint64[n+1])
            qres = np.ascontiguousarray(queries[0], np.uint8)
            qoff = np.ascontiguousarray(queries[1], np.uint64)
            qlens = np.diff(qoff.astype(np.int64))
        else:
            qres = np.ascontiguousarray(np.concatenate(queries), np.uint8)
            qlens = np.array([len(q) for q in queries], np.int64)
            qoff = np.zeros(len(qlens) + 1, np.uint64)
            qoff[1:] = np.cumsum(qlens)
        ta = np.zeros(len(tasks), NUCL_TASK_DTYPE)
        tarr = np.asarray(tasks, np.int64).reshape(-1, 3)
        ta["query"], ta["target"], ta["diagonal"] = tarr[:, 0], tarr[:, 1], tarr[:, 2] & 0xffff
        slots = (2 * qlens[tarr[:, 0]] + 72).astype(np.uint64)
        coff = np.zeros(len(ta) + 1, np.uint64)
        coff[1:] = np.cumsum(slots)
        cig = np.empty(int(coff[-1]) + 1, np.uint32)   # only out[i].n_cigar entries of each slot are written
        out = np.zeros(len(ta), NUCL_ALN_DTYPE)
        self._check(self.lib.b200_nucl_align(self.h, _p(qres), _p(qoff), ctypes.c_uint32(len(qlens)), _p(ta), _u64(len(ta)), go, ge,
                                             zdrop, _p(out), _p(cig), _p(coff)))
        if not decode:
            return out, (cig, coff), None
        cigars = [cig[int(coff[i]):int(coff[i]) + int(out["n_cigar"][i])].copy() for i in range(len(ta))]
        bts = ["".join("MID"[int(c & 0xf)] * int(c >> 4) for c in cg) for cg in cigars]
        return out, cigars, bts

    def sw_job(self, queries, pairs, go=11, ge=1):
        cq = _cqueries(queries)
        pa = pairs if getattr(pairs, "dtype", None) == PAIR_DTYPE else self._pairs(pairs)
        h = _vp()
        self._check(self.lib.b200_sw_job_create(self.h, cq, len(queries), _p(pa), _u64(len(pa)), go, ge, ctypes.byref(h)))
        return Job(self, h, "sw", n=len(pa))


class MultiContext:
    """Several GPUs of a node behind one handle (include/b200_multi.h): query-sharded (DB replicated) or target-sharded."""

    def __init__(self, devices=None):
        self.lib = load_library()
        h = _vp()
        if devices is None:
            rc = self.lib.b200_multi_create(None, 0, ctypes.byref(h))
        else:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self.lib.b200_multi_create(arr, len(devices), ctypes.byref(h))
        if rc != 0:
            raise B200Error("b200_multi_create failed with %d" % rc)
        self.h = h
        self.lib.b200_multi_last_error.restype = ctypes.c_char_p
        self.n_seq = 0

    @property
    def size(self):
        return int(self.lib.b200_multi_size(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise B200Error("b200_multi error %d: %s" % (rc, self.lib.b200_multi_last_error(self.h).decode()))

    def load_db(self, residues, offsets, alphabet, shard_targets=False):
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        self._check(self.lib.b200_multi_db_load(self.h, _p(residues), _p(offsets), _u64(len(offsets) - 1), int(alphabet), 1 if shard_targets else 0))
        self.n_seq = len(offsets) - 1
        self._db_len = np.diff(offsets.astype(np.int64))

    def db_lengths(self):
        return self._db_len

    def ungapped_scan(self, queries, min_score_excl=15, max_hits=300):
        cq = _cqueries(queries)
        nq = len(queries)
        hits = np.zeros((nq, max_hits), HIT_DTYPE)
        n_hits = np.zeros(nq, np.uint32)
        self._check(self.lib.b200_multi_ungapped_scan(self.h, cq, nq, int(min_score_excl), ctypes.c_uint32(max_hits), _p(hits), _p(n_hits)))
        return hits, n_hits
