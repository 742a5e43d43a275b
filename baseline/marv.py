"""baseline/marv.py -- BASELINE / measurement infrastructure (never imported by mmseqs2_b200/).

ctypes driver for baseline/_ref/libmarv_harness.so: the reference's own GPU scorer `class Marv`
(lib/libmarv/src/marv.h:6-58), compiled in place by baseline/Makefile, run on the same target DB and queries as
libb200align.so so that bench.py can print the reference's GPU GCUPS next to ours on the same B200.

The target DB is laid out as `makepaddedseqdb` writes it (src/util/makepaddedseqdb.cpp:20,66-98): sequences sorted by
length ascending, residues as numeric codes, each padded to a multiple of 4 with code 20; offsets[n+1] / lengths[n] as
ungappedprefilter.cpp:127-139 derives them from the index.
"""
import ctypes
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libmarv_harness.so")

TABLES = {"as_shipped": 0, "sm89_half2": 89, "sm90_dpx": 90, "sm103_dpx": 103}


def available():
    return os.path.exists(SO)


def padded_db(res, off):
    """-> (bytes uint8, offsets uint64[n+1], lengths int32[n], order) in makepaddedseqdb layout"""
    off = np.asarray(off, np.int64)
    lens = (off[1:] - off[:-1]).astype(np.int64)
    order = np.argsort(lens, kind="stable")
    sl = lens[order]
    pl = (sl + 3) // 4 * 4
    poff = np.zeros(len(sl) + 1, np.uint64)
    poff[1:] = np.cumsum(pl)
    out = np.full(int(poff[-1]), 20, np.uint8)
    # vectorised scatter: destination index of every residue
    src_start = off[:-1][order]
    dst_start = poff[:-1].astype(np.int64)
    total = int(sl.sum())
    seq_of = np.repeat(np.arange(len(sl)), sl)
    within = np.arange(total) - np.repeat(np.cumsum(sl) - sl, sl)
    out[dst_start[seq_of] + within] = res[src_start[seq_of] + within]
    offsets = poff.copy()
    offsets[-1] = offsets[-2] + np.uint64(sl[-1])      # ungappedprefilter.cpp:136: offsets.back() + lengths.back()
    return out, offsets, sl.astype(np.int32), order


class Marv:
    def __init__(self, res, off, max_seqs=300, alignment_type=0):
        if not available():
            raise RuntimeError("baseline/_ref/libmarv_harness.so not built (make -C baseline)")
        L = ctypes.CDLL(SO)
        L.marvh_create.restype = ctypes.c_void_p
        L.marvh_create.argtypes = [ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int]
        L.marvh_load_db.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 3 + [ctypes.c_size_t]
        L.marvh_set_table.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.marvh_scan.restype = ctypes.c_long
        L.marvh_scan.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t] + [ctypes.c_void_p] * 6
        L.marvh_destroy.argtypes = [ctypes.c_void_p]
        self.L = L
        self.data, self.offsets, self.lengths, self.order = padded_db(res, off)
        self.max_seqs = max_seqs
        self.h = L.marvh_create(len(self.lengths), 21, int(self.lengths[-1]), max_seqs, alignment_type)
        rc = L.marvh_load_db(self.h, self.data.ctypes.data, self.offsets.ctypes.data, self.lengths.ctypes.data, self.data.nbytes)
        if rc:
            raise RuntimeError("marvh_load_db failed")
        self.ids = np.zeros(max_seqs, np.uint32)
        self.scores = np.zeros(max_seqs, np.int32)
        self.qend = np.zeros(max_seqs, np.int32)
        self.dbend = np.zeros(max_seqs, np.int32)
        self.stats = np.zeros(3, np.float64)

    def set_table(self, name):
        if self.L.marvh_set_table(self.h, TABLES[name]):
            raise RuntimeError("marvh_set_table")

    def scan(self, qseq, profile):
        """one Marv::scan; -> (original target ids, scores, libmarv's own (seconds, gcups, overflows))"""
        q = np.ascontiguousarray(qseq, np.uint8)
        p = np.ascontiguousarray(profile, np.int8)
        n = self.L.marvh_scan(self.h, q.ctypes.data, len(q), p.ctypes.data, self.ids.ctypes.data, self.scores.ctypes.data,
                              self.qend.ctypes.data, self.dbend.ctypes.data, self.stats.ctypes.data)
        return self.order[self.ids[:n]], self.scores[:n].copy(), tuple(self.stats)

    def close(self):
        if self.h:
            self.L.marvh_destroy(self.h)
            self.h = None


def time_tables(res, off, qseqs, profiles, tables=("as_shipped", "sm90_dpx", "sm103_dpx"), max_seqs=300, warm=2):
    """Marv::scan over the given queries, one call per query as ungappedprefilter.cpp:165-207 does, for each kernel table.
    GCUPS by wall clock around the calls (what a caller of Marv sees) and by libmarv's own per-scan Stats."""
    m = Marv(res, off, max_seqs)
    db_res = float(np.asarray(off, np.int64)[-1])
    out, hits = {}, {}
    for tb in tables:
        m.set_table(tb)
        for i in range(min(warm, len(qseqs))):
            m.scan(qseqs[i], profiles[i])
        cells, own_s = 0.0, 0.0
        lists = []
        t0 = time.perf_counter()
        for q, p in zip(qseqs, profiles):
            ids, sc, st = m.scan(q, p)
            cells += len(q) * db_res
            own_s += st[0]
            lists.append((ids, sc))
        dt = time.perf_counter() - t0
        out[tb] = {"gcups_wall": cells / 1e9 / dt, "gcups_own_stats": cells / 1e9 / own_s if own_s > 0 else None,
                   "queries": len(qseqs), "ms_per_query": 1e3 * dt / len(qseqs)}
        hits[tb] = lists
    m.close()
    return out, hits
