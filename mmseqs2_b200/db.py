"""ctypes binding of include/b200_db.h: the DB triple (data / .index / .dbtype, plain or zstd-compressed) as DBReader / DBWriter define
it, the letter mapping of Sequence::mapSequence, and the reference's modules over DB files:

  mmseqs ungappedprefilter   prefilter_db()          mmseqs align              align_db()
  mmseqs rescorediagonal     rescorediagonal_db()    mmseqs makepaddedseqdb    make_padded_db()  (host only, with the repeat masker)

Host-side formats; the DPs and scorers run in libb200align.so."""
import ctypes

import numpy as np

from .api import B200Error, _p, _u64, _vp, load_library

DBTYPE_AMINO_ACIDS, DBTYPE_NUCLEOTIDES, DBTYPE_HMM_PROFILE, DBTYPE_ALIGNMENT_RES, DBTYPE_PREFILTER_RES = 0, 1, 2, 5, 7


def _lib():
    lib = load_library()
    lib.b200h_db_last_error.restype = ctypes.c_char_p
    lib.b200h_db_size.restype = ctypes.c_uint64
    lib.b200h_db_key.restype = ctypes.c_uint32
    lib.b200h_db_data.restype = ctypes.c_void_p
    lib.b200h_db_entry_len.restype = ctypes.c_uint64
    lib.b200h_db_id.restype = ctypes.c_int64
    for n in ("b200h_db_size", "b200h_db_type", "b200h_db_close"):
        getattr(lib, n).argtypes = [_vp]
    lib.b200h_db_key.argtypes = [_vp, ctypes.c_uint64]
    lib.b200h_db_data.argtypes = [_vp, ctypes.c_uint64]
    lib.b200h_db_entry_len.argtypes = [_vp, ctypes.c_uint64]
    lib.b200h_db_id.argtypes = [_vp, ctypes.c_uint32]
    return lib


class DB:
    """DBReader: entries by id = rank of the key."""

    def __init__(self, path):
        self.lib = _lib()
        self.h = _vp()
        if self.lib.b200h_db_open(path.encode(), ctypes.byref(self.h)) != 0:
            raise B200Error(self.lib.b200h_db_last_error().decode())

    def __len__(self):
        return int(self.lib.b200h_db_size(self.h))

    @property
    def dbtype(self):
        return int(self.lib.b200h_db_type(self.h))

    def key(self, i):
        return int(self.lib.b200h_db_key(self.h, i))

    def entry_len(self, i):
        return int(self.lib.b200h_db_entry_len(self.h, i))

    def data(self, i):
        """payload of entry i without its NUL terminator"""
        return ctypes.string_at(self.lib.b200h_db_data(self.h, i), self.entry_len(i) - 1)

    def id_of(self, key):
        return int(self.lib.b200h_db_id(self.h, key))

    def close(self):
        if self.h:
            self.lib.b200h_db_close(self.h)
            self.h = None


def write_db(path, dbtype, keys, entries):
    """DBWriter with one thread: entries (bytes) under keys, in the given order."""
    lib = _lib()
    w = _vp()
    if lib.b200h_dbw_open(path.encode(), int(dbtype), ctypes.byref(w)) != 0:
        raise B200Error(lib.b200h_db_last_error().decode())
    for k, e in zip(keys, entries):
        if lib.b200h_dbw_write(w, ctypes.c_uint32(int(k)), e, _u64(len(e))) != 0:
            raise B200Error(lib.b200h_db_last_error().decode())
    if lib.b200h_dbw_close(w) != 0:
        raise B200Error(lib.b200h_db_last_error().decode())


def aa2num_table(alphabet, nucleotide=False):
    """BaseMatrix::aa2num after setupLetterMapping for a matrix alphabet given in numeric order (bytes / str)."""
    lib = _lib()
    a = alphabet.encode() if isinstance(alphabet, str) else bytes(alphabet)
    t = np.zeros(256, np.uint8)
    lib.b200h_aa2num_table(a, len(a), 1 if nucleotide else 0, _p(t))
    return t


def align_db(ctx, submat, alphabet, query_db, target_db, prefilter_db, alignment_db, params, evalue=None, add_backtrace=True,
             bucket_queries=4096):
    """`mmseqs align` over DB files -> (number of alignments computed, number of records written)"""
    lib = _lib()
    a = alphabet.encode() if isinstance(alphabet, str) else bytes(alphabet)
    mat = np.ascontiguousarray(submat.mat, np.int16)
    pb = np.ascontiguousarray(submat.pback, np.float64)
    na, nr = ctypes.c_uint64(0), ctypes.c_uint64(0)
    rc = lib.b200_align_db(ctx.h, query_db.encode(), target_db.encode(), prefilter_db.encode(), alignment_db.encode(), _p(mat), _p(pb), a,
                           int(submat.A), ctypes.byref(params), None if evalue is None else ctypes.byref(evalue),
                           1 if add_backtrace else 0, ctypes.c_uint32(bucket_queries), ctypes.byref(na), ctypes.byref(nr))
    ctx._check(rc)
    return int(na.value), int(nr.value)


def prefilter_db(ctx, submat, alphabet, query_db, target_db, prefilter_db_path, comp_bias=True, comp_bias_scale=1.0, min_diag_score=15,
                 max_res_list_len=300, bucket_queries=64):
    """`mmseqs ungappedprefilter` over DB files -> number of hits written"""
    lib = _lib()
    a = alphabet.encode() if isinstance(alphabet, str) else bytes(alphabet)
    mat = np.ascontiguousarray(submat.mat, np.int16)
    pb = np.ascontiguousarray(submat.pback, np.float64)
    nh = ctypes.c_uint64(0)
    rc = lib.b200_prefilter_db(ctx.h, query_db.encode(), target_db.encode(), prefilter_db_path.encode(), _p(mat), _p(pb), a, int(submat.A),
                               1 if comp_bias else 0, ctypes.c_float(comp_bias_scale), int(min_diag_score), ctypes.c_uint32(max_res_list_len),
                               ctypes.c_uint32(bucket_queries), ctypes.byref(nh))
    ctx._check(rc)
    return int(nh.value)


def make_padded_db(src_db, dst_db, alphabet, likelihood_ratio=None, mask=1, mask_prob=0.9, mask_lower_case=0, mask_n_repeat=0,
                   write_lookup=1, threads=0):
    """`mmseqs makepaddedseqdb` (same parameter names and defaults): the padded GPU sequence DB of an amino-acid DB.
    likelihood_ratio: [A, A] doubles, the reference's ProbabilityMatrix (needed when mask != 0)."""
    import os
    lib = _lib()
    lib.b200h_paddeddb_last_error.restype = ctypes.c_char_p
    table = aa2num_table(alphabet)
    a = alphabet.encode() if isinstance(alphabet, str) else bytes(alphabet)
    lr = None if likelihood_ratio is None else np.ascontiguousarray(likelihood_ratio, np.float64).reshape(-1)
    if lr is not None and lr.size != len(a) * len(a):
        raise ValueError("likelihood_ratio must be [A, A]")
    if threads <= 0:
        threads = len(os.sched_getaffinity(0))
    rc = lib.b200h_make_padded_db(src_db.encode(), dst_db.encode(), _p(table), len(a), None if lr is None else _p(lr), int(mask),
                                  ctypes.c_double(mask_prob), int(mask_lower_case), int(mask_n_repeat), int(write_lookup), int(threads))
    if rc != 0:
        raise B200Error(lib.b200h_paddeddb_last_error().decode())


def tantan_probabilities(codes, likelihood_ratio):
    """posterior probability of "inside a repeat" per residue (tantan as Masker::maskSequence configures it)"""
    lib = _lib()
    s = np.ascontiguousarray(codes, np.uint8)
    lr = np.ascontiguousarray(likelihood_ratio, np.float64)
    out = np.zeros(len(s), np.float32)
    if lib.b200h_tantan_probabilities(_p(s), len(s), int(lr.shape[0]), _p(lr), _p(out)) != 0:
        raise B200Error("b200h_tantan_probabilities: bad arguments")
    return out


def mask_sequence(codes, text, likelihood_ratio, mask_tantan=True, mask_prob=0.9, mask_lower_case=False, mask_n_repeat=0):
    """Masker::maskSequence on numeric codes -> (masked codes, number of masked residues)"""
    lib = _lib()
    s = np.array(codes, np.uint8)
    lr = np.ascontiguousarray(likelihood_ratio, np.float64)
    n = lib.b200h_mask_sequence(_p(s), text, len(s), int(lr.shape[0]), _p(lr), 1 if mask_tantan else 0, ctypes.c_double(mask_prob),
                                1 if mask_lower_case else 0, int(mask_n_repeat))
    if n < 0:
        raise B200Error("b200h_mask_sequence: bad arguments")
    return s, int(n)


class RescoreParams(ctypes.Structure):
    """b200_rescore_params: the rescorediagonal parameters (same meaning and defaults as the module's command line)"""
    _fields_ = [("rescore_mode", ctypes.c_int), ("eval_thr", ctypes.c_double), ("cov_thr", ctypes.c_float), ("cov_mode", ctypes.c_int),
                ("seq_id_thr", ctypes.c_float), ("aln_len_thr", ctypes.c_int), ("seq_id_mode", ctypes.c_int), ("include_identity", ctypes.c_int),
                ("add_backtrace", ctypes.c_int), ("sort_results", ctypes.c_int)]

    def __init__(self, rescore_mode=0, eval_thr=1e-3, cov_thr=0.0, cov_mode=0, seq_id_thr=0.0, aln_len_thr=0, seq_id_mode=0, include_identity=0,
                 add_backtrace=0, sort_results=0):
        super().__init__(rescore_mode, eval_thr, cov_thr, cov_mode, seq_id_thr, aln_len_thr, seq_id_mode, include_identity, add_backtrace, sort_results)


RESCORE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p)


def ascii_matrix(submat, alphabet, nucleotide=False):
    """SubstitutionMatrix::createAsciiSubMat: [123, 123] int8"""
    lib = _lib()
    a = alphabet.encode() if isinstance(alphabet, str) else bytes(alphabet)
    mat = np.ascontiguousarray(submat.mat, np.int16)
    out = np.zeros((123, 123), np.int8)
    lib.b200h_ascii_matrix(_p(mat), a, len(a), 1 if nucleotide else 0, _p(out))
    return out


def rescorediagonal_db(ctx, submat, alphabet, query_db, target_db, prefilter_db, out_db, params, evalue=None, bucket_queries=4096, scorer=None):
    """`mmseqs rescorediagonal` over DB files -> (hits scored, records written).  scorer: None = the device scorer of ctx; a RESCORE_FN
    runs the module's host half with that scorer instead (tests/test_rescore_module.py checks the host half this way without a GPU)."""
    lib = _lib()
    lib.b200h_rescore_module_last_error.restype = ctypes.c_char_p
    a = alphabet.encode() if isinstance(alphabet, str) else bytes(alphabet)
    mat = np.ascontiguousarray(submat.mat, np.int16)
    nh, nr = ctypes.c_uint64(0), ctypes.c_uint64(0)
    ev = None if evalue is None else ctypes.byref(evalue)
    if scorer is None:
        rc = lib.b200_rescorediagonal_db(ctx.h, query_db.encode(), target_db.encode(), prefilter_db.encode(), out_db.encode(), _p(mat), a, len(a),
                                         ctypes.byref(params), ev, ctypes.c_uint32(bucket_queries), ctypes.byref(nh), ctypes.byref(nr))
        ctx._check(rc)
    else:
        rc = lib.b200h_rescorediagonal_db_with(scorer, None, query_db.encode(), target_db.encode(), prefilter_db.encode(), out_db.encode(), _p(mat), a,
                                               len(a), ctypes.byref(params), ev, ctypes.c_uint32(bucket_queries), ctypes.byref(nh), ctypes.byref(nr))
        if rc != 0:
            raise B200Error(lib.b200h_rescore_module_last_error().decode())
    return int(nh.value), int(nr.value)
