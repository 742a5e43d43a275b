"""ctypes binding of include/b200_alignment.h: the batched `align` step (Alignment::run, src/alignment/Alignment.cpp:312-420),
its E-value statistics (EvalueComputation.h) and the prefilter / alignment record formats (QueryMatcher.h:78-132,
Matcher.cpp:282-329).  Host-side mirror only: every DP of the batch runs in libb200align.so."""
import ctypes

import numpy as np

from .api import B200Error, _p, _u64, _vp, load_library

PREF_HIT_DTYPE = np.dtype([("seq_id", np.uint32), ("pref_score", np.int32), ("diagonal", np.uint16), ("pad_", np.uint16)])
RESULT_DTYPE = np.dtype([("db_key", np.uint32), ("score", np.int32), ("qcov", np.float32), ("dbcov", np.float32),
                         ("seq_id", np.float32), ("pad0_", np.uint32), ("eval", np.float64), ("aln_length", np.uint32),
                         ("q_start", np.int32), ("q_end", np.int32), ("q_len", np.int32), ("db_start", np.int32),
                         ("db_end", np.int32), ("db_len", np.int32), ("pad1_", np.uint32), ("bt_off", np.uint64),
                         ("bt_len", np.uint32), ("pad_", np.uint32)])

# Matcher::SCORE_ONLY / SCORE_COV / SCORE_COV_SEQID (Matcher.h:24-26)
SCORE_ONLY, SCORE_COV, SCORE_COV_SEQID = 0, 1, 2
UINT32_MAX = 0xFFFFFFFF


class EvalueParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ("lambda_", "K", "a_J", "b_J", "a_I", "b_I", "alpha_J", "beta_J", "alpha_I", "beta_I",
                                               "sigma", "tau")] + [("db_residues", ctypes.c_uint64)]

    @classmethod
    def defaults(cls, matrix, gap_open, gap_extend, db_residues, gapped=True):
        """The hard-coded sets of EvalueComputation.h:56-81 ("blosum62.out" 11/1, "nucleotide.out" 5/2 and 7/1, ungapped blosum62)."""
        lib = load_library()
        p = cls()
        rc = lib.b200h_evalue_defaults(matrix.encode(), gap_open, gap_extend, 1 if gapped else 0, _u64(db_residues), ctypes.byref(p))
        if rc != 0:
            raise B200Error("no built-in Gumbel parameters for %s %d/%d (the reference would run ALP; pass evaluer.parameters())"
                            % (matrix, gap_open, gap_extend))
        return p

    def evalue(self, score, query_len):
        lib = load_library()
        lib.b200h_evalue.restype = ctypes.c_double
        return float(lib.b200h_evalue(ctypes.byref(self), ctypes.c_double(score), ctypes.c_double(query_len)))

    def bit_score(self, score):
        lib = load_library()
        lib.b200h_bit_score.restype = ctypes.c_double
        return float(lib.b200h_bit_score(ctypes.byref(self), ctypes.c_double(score)))


class AlignParams(ctypes.Structure):
    """Parameters of `mmseqs align` that reach Alignment::run's per-query loop; defaults = the reference's defaults."""
    _fields_ = [("gap_open", ctypes.c_int), ("gap_extend", ctypes.c_int), ("sw_mode", ctypes.c_int), ("eval_thr", ctypes.c_double),
                ("cov_thr", ctypes.c_float), ("cov_mode", ctypes.c_int), ("seq_id_thr", ctypes.c_float), ("aln_len_thr", ctypes.c_int),
                ("seq_id_mode", ctypes.c_int), ("max_accept", ctypes.c_uint32), ("max_rejected", ctypes.c_uint32),
                ("comp_bias", ctypes.c_int), ("comp_bias_scale", ctypes.c_float), ("include_identity", ctypes.c_int)]

    def __init__(self, gap_open=11, gap_extend=1, sw_mode=SCORE_COV_SEQID, eval_thr=1e-3, cov_thr=0.0, cov_mode=0, seq_id_thr=0.0,
                 aln_len_thr=0, seq_id_mode=0, max_accept=0x7FFFFFFF, max_rejected=0x7FFFFFFF, comp_bias=True, comp_bias_scale=1.0,
                 include_identity=False):
        super().__init__(gap_open, gap_extend, sw_mode, eval_thr, cov_thr, cov_mode, seq_id_thr, aln_len_thr, seq_id_mode, max_accept,
                         max_rejected, 1 if comp_bias else 0, comp_bias_scale, 1 if include_identity else 0)


def parse_prefilter_hits(entry):
    """QueryMatcher::parsePrefilterHits on one entry (bytes without the trailing NUL)."""
    lib = load_library()
    lib.b200h_parse_prefilter_hits.restype = ctypes.c_size_t
    cap = entry.count(b"\n") + 2
    out = np.zeros(cap, PREF_HIT_DTYPE)
    n = lib.b200h_parse_prefilter_hits(entry + b"\0", _p(out), ctypes.c_size_t(cap))
    return out[:n]


def prefilter_hits_to_buffer(hits):
    """QueryMatcher::prefilterHitToBuffer for every hit, concatenated."""
    lib = load_library()
    lib.b200h_prefilter_hit_to_buffer.restype = ctypes.c_size_t
    hits = np.ascontiguousarray(hits, PREF_HIT_DTYPE)
    buf = ctypes.create_string_buffer(48)
    parts = []
    for i in range(len(hits)):
        n = lib.b200h_prefilter_hit_to_buffer(buf, _vp(hits.ctypes.data + i * PREF_HIT_DTYPE.itemsize))
        parts.append(buf.raw[:n])
    return b"".join(parts)


def result_to_buffer(result, backtrace=b"", add_backtrace=False, compress=True):
    """Matcher::resultToBuffer for one RESULT_DTYPE record."""
    lib = load_library()
    lib.b200h_result_to_buffer.restype = ctypes.c_size_t
    r = np.zeros(1, RESULT_DTYPE)
    r[0] = result
    r["bt_len"][0] = len(backtrace)
    buf = ctypes.create_string_buffer(512 + 2 * len(backtrace))
    n = lib.b200h_result_to_buffer(buf, _p(r), backtrace, 1 if add_backtrace else 0, 1 if compress else 0)
    return buf.raw[:n]


def align_batch(ctx, submat, queries, hit_lists, params, evalue, query_keys=None, target_keys=None):
    """Alignment::run for a batch.  queries: list of numeric uint8 arrays; hit_lists: per query the DB-local target ids in
    prefilter-list order.  -> (per-query list of RESULT_DTYPE arrays, backtrace pool bytes, number of alignments computed)"""
    lib = load_library()
    nq = len(queries)
    qoff = np.zeros(nq + 1, np.uint64)
    qoff[1:] = np.cumsum([len(q) for q in queries])
    qres = np.concatenate([np.ascontiguousarray(q, np.uint8) for q in queries]) if nq else np.zeros(0, np.uint8)
    hoff = np.zeros(nq + 1, np.uint64)
    hoff[1:] = np.cumsum([len(h) for h in hit_lists])
    n_hits = int(hoff[-1])
    htg = (np.concatenate([np.asarray(h, np.uint32) for h in hit_lists]) if n_hits else np.zeros(0, np.uint32)).astype(np.uint32)
    htg = np.ascontiguousarray(htg)
    qk = None if query_keys is None else np.ascontiguousarray(query_keys, np.uint32)
    tk = None if target_keys is None else np.ascontiguousarray(target_keys, np.uint32)
    res = np.zeros(max(1, n_hits), RESULT_DTYPE)
    nres = np.zeros(max(1, nq), np.uint32)
    mat = np.ascontiguousarray(submat.mat, np.int16)
    pb = np.ascontiguousarray(submat.pback, np.float64)
    n_aln = ctypes.c_uint64(0)
    # worst case of the backtrace pool: qlen + tlen per accepted hit
    bt_cap = 16
    if params.sw_mode == SCORE_COV_SEQID or params.include_identity:   # scoreIdentical leaves an all-M backtrace in every mode
        lens = ctx.db_lengths()
        for qi in range(nq):
            if len(hit_lists[qi]):
                ids = np.minimum(np.asarray(hit_lists[qi], np.int64), len(lens) - 1)   # range errors are the library's to report
                bt_cap += int(len(queries[qi]) * len(hit_lists[qi]) + lens[ids].sum())
    pool = np.empty(bt_cap, np.uint8)            # written by the library up to the last accepted backtrace; never read beyond
    fn = lib.b200_multi_align_batch if type(ctx).__name__ == "MultiContext" else lib.b200_align_batch   # same arguments, several GPUs
    rc = fn(ctx.h, _p(mat), _p(pb), int(submat.A), _p(qres), _p(qoff), _p(qk), ctypes.c_uint32(nq), _p(hoff), _p(htg),
            _p(tk), ctypes.byref(params), ctypes.byref(evalue), _p(res), _p(nres), _p(pool), _u64(bt_cap),
            ctypes.byref(n_aln))
    ctx._check(rc)
    out = [res[int(hoff[i]):int(hoff[i]) + int(nres[i])] for i in range(nq)]
    return out, pool, int(n_aln.value)


def align_batch_nucl(ctx, reads, hit_lists, hit_diagonals, hit_reverse, params, evalue, zdrop=40, query_keys=None, target_keys=None):
    """Alignment::run for a nucleotide search batch.  reads: list of numeric uint8 arrays (A,C,T,G,X = 0..4); per read the
    DB-local target ids, prefilter diagonals (int16) and strand flags (or None) in list order."""
    lib = load_library()
    nq = len(reads)
    qoff = np.zeros(nq + 1, np.uint64)
    qoff[1:] = np.cumsum([len(q) for q in reads])
    qres = np.concatenate([np.ascontiguousarray(q, np.uint8) for q in reads]) if nq else np.zeros(0, np.uint8)
    hoff = np.zeros(nq + 1, np.uint64)
    hoff[1:] = np.cumsum([len(h) for h in hit_lists])
    n_hits = int(hoff[-1])
    cat = lambda xs, dt: np.ascontiguousarray(np.concatenate([np.asarray(x, dt) for x in xs]) if n_hits else np.zeros(0, dt), dt)  # noqa: E731
    htg = cat(hit_lists, np.uint32)
    hdg = cat(hit_diagonals, np.int16)
    hrv = None if hit_reverse is None else cat(hit_reverse, np.uint8)
    qk = None if query_keys is None else np.ascontiguousarray(query_keys, np.uint32)
    tk = None if target_keys is None else np.ascontiguousarray(target_keys, np.uint32)
    res = np.zeros(max(1, n_hits), RESULT_DTYPE)
    nres = np.zeros(max(1, nq), np.uint32)
    n_aln = ctypes.c_uint64(0)
    bt_cap = 16 + sum((2 * len(reads[i]) + 72) * len(hit_lists[i]) for i in range(nq))
    pool = np.empty(bt_cap, np.uint8)
    rc = lib.b200_align_batch_nucl(ctx.h, _p(qres), _p(qoff), _p(qk), ctypes.c_uint32(nq), _p(hoff), _p(htg), _p(hdg), _p(hrv), _p(tk),
                                   ctypes.byref(params), int(zdrop), ctypes.byref(evalue), _p(res), _p(nres), _p(pool), _u64(bt_cap),
                                   ctypes.byref(n_aln))
    ctx._check(rc)
    out = [res[int(hoff[i]):int(hoff[i]) + int(nres[i])] for i in range(nq)]
    return out, pool, int(n_aln.value)


def records(results, pool, add_backtrace=True, compress=True):
    """The alignment DB entry of one query (what Alignment::run writes for it, Alignment.cpp:505-512)."""
    parts = []
    for r in results:
        bt = bytes(pool[int(r["bt_off"]):int(r["bt_off"]) + int(r["bt_len"])])
        parts.append(result_to_buffer(r, bt, add_backtrace, compress))
    return b"".join(parts)
