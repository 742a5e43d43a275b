"""tests/golden/make_align_golden.py -- fixtures for the batched align step (include/b200_alignment.h) from the REFERENCE's own
Matcher / EvalueComputation / QueryMatcher code (oracle/_ref, i.e. needs /root/reference; build container only):
    python tests/golden/make_align_golden.py
Writes tests/golden/align_v1.npz:
  * E-values and bit scores of EvalueComputation (blosum62 11/1) on a (score, query length, DB size) grid
  * alignment records of Matcher::resultToBuffer for hand-picked field values (format corner cases)
  * prefilter entries parsed and re-serialised by QueryMatcher
  * whole alignment-DB entries (text) of 40 real-protein queries (examples_v1.npz) against their prefilter lists for several
    parameter sets of `mmseqs align` -- what Alignment::run writes per query
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.pyoracle import Ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

from align_configs import CONFIGS, NUCL_CONFIGS  # noqa: E402  (shared with tests/test_alignment_batch.py)


def hit_lists(ex, n_per_query=120):
    """prefilter-like lists: the best ungapped scores first (ties by id), a few low scorers at the end"""
    lists = []
    ung = ex["ungapped"].astype(np.int32)
    for qi in range(ung.shape[0]):
        order = np.lexsort((np.arange(ung.shape[1]), -ung[qi]))
        lists.append(np.concatenate([order[:n_per_query], order[-5:]]).astype(np.uint32))
    return lists


def main():
    ref = Ref()
    out = {}
    # ---- E-values ----------------------------------------------------------------------------------------------------
    grid = [(s, ql, db) for s in (1, 17, 40, 41, 64, 100, 254, 255, 1000, 32767) for ql in (1, 7, 50, 350, 2000, 65535)
            for db in (1, 10000, 359372871, 7000000000)]
    ev = np.zeros((len(grid), 5), np.float64)
    for i, (s, ql, db) in enumerate(grid):
        bits = ref.bit_score(11, 1, db, s)
        ev[i] = (s, ql, db, ref.evalue(11, 1, db, s, ql), bits)
    out["evalue_grid"] = ev
    # ---- record formatting -------------------------------------------------------------------------------------------
    cases = [
        (7, 250, 1.0, 1.5e-70, 0, 99, 100, 0, 99, 100, b"M" * 100),
        (4000000000, 31, 0.0999, 9.87e-4, 3, 40, 350, 12, 51, 77, b"MMMIMMDDMM"),
        (12, 5, 0.005, 12.5, -1, 10, 20, -1, 30, 40, b""),
        (0, -3, 0.0, 0.0, 0, 0, 1, 0, 0, 1, b"M"),
        (99, 1234, 0.4567, 1e-300, 100, 2000, 2001, 5, 1900, 30000, b"M" * 7 + b"I" * 12 + b"M" * 3 + b"D"),
        (5, 77, 0.9999, 3.0e5, 1, 2, 3, 4, 5, 6, b"DIM"),
        (6, 77, 0.01, 1.0, 1, 2, 3, 4, 5, 6, b"IIII"),
        (8, 77, 0.1, 0.99951, 1, 2, 3, 4, 5, 6, b"MMMM"),
    ]
    rec = []
    for c in cases:
        for add_bt, comp in ((False, True), (True, True), (True, False)):
            rec.append(ref.result_to_buffer(*c[:10], backtrace=c[10], add_backtrace=add_bt, compress=comp))
    out["record_cases"] = np.array([(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9]) for c in cases], np.float64)
    out["record_case_bt"] = np.array([c[10] for c in cases])
    out["record_case_text"] = np.array(rec)
    # ---- prefilter entries -----------------------------------------------------------------------------------------------
    entry = b"17\t250\t0\n4000000000\t-12\t-5\n3\t15\t32767\n8\t1\t-32768\n9 40 12\n"
    ids, sc, dg, txt = ref.prefilter_roundtrip(entry)
    out["pref_entry"] = np.array(entry)
    out["pref_ids"], out["pref_scores"], out["pref_diags"], out["pref_text"] = ids, sc, dg, np.array(txt)
    # ---- alignment-DB entries ----------------------------------------------------------------------------------------------
    ex = np.load(os.path.join(HERE, "examples_v1.npz"))
    td, to = ex["tdata"], ex["toff"].astype(np.int64)
    qd, qo = ex["qdata"], ex["qoff"]
    n_t = len(to) - 1
    tkeys = (1000 + 3 * np.arange(n_t)).astype(np.uint32)
    lists = hit_lists(ex)
    out["hit_off"] = np.cumsum([0] + [len(h) for h in lists]).astype(np.uint64)
    out["hit_targets"] = np.concatenate(lists)
    out["target_keys"] = tkeys
    db_res = int(to[-1])
    out["db_residues"] = np.array(db_res)
    for name, kw in CONFIGS.items():
        texts, nal = [], []
        for qi in range(len(qo) - 1):
            q = qd[int(qo[qi]):int(qo[qi + 1])]
            txt, na, nacc = ref.align_query(q, 5000 + qi, td, to, lists[qi], tkeys[lists[qi]], db_res, **kw)
            texts.append(txt); nal.append(na)
        out["cfg_%s_text" % name] = np.array(texts)
        out["cfg_%s_naligned" % name] = np.array(nal, np.int64)
        print(name, "records:", sum(t.count(b"\n") for t in texts), "alignments:", sum(nal))
    # ---- identity hits: the queries themselves are part of the DB, same key => isIdentity --------------------------------
    qs = [qd[int(qo[i]):int(qo[i + 1])] for i in range(12)]
    ts = [td[int(to[i]):int(to[i + 1])] for i in range(60)]
    seqs = qs + ts
    sd = np.concatenate(seqs); so = np.cumsum([0] + [len(s) for s in seqs]).astype(np.int64)
    skeys = np.arange(len(seqs), dtype=np.uint32) * 2 + 10
    out["self_data"], out["self_off"], out["self_keys"] = sd, so, skeys
    for mode in (0, 1, 2):
        texts = []
        for qi in range(len(qs)):
            hl = np.array([qi] + list(range(12, 72)) + [(qi + 1) % 12], np.uint32)
            txt, _, _ = ref.align_query(qs[qi], int(skeys[qi]), sd, so, hl, skeys[hl], int(so[-1]), sw_mode=mode, eval_thr=1e-3,
                                        include_identity=True, add_backtrace=(mode == 2))
            texts.append(txt)
        out["self_mode%d_text" % mode] = np.array(texts)
    # ---- nucleotide search: reads of nucl_v1.npz, their true hit first, decoys behind it, every third read on the reverse strand ----
    nv = np.load(os.path.join(HERE, "nucl_v1.npz"))
    ntd, nto = nv["tdata"], nv["toff"].astype(np.int64)
    nqd, nqo, ntasks = nv["qdata"], nv["qoff"], nv["tasks"]
    rng = np.random.default_rng(99)
    comp = np.array([2, 3, 0, 1, 4], np.uint8)
    n_t = len(nto) - 1
    nkeys = (7 + 5 * np.arange(n_t)).astype(np.uint32)
    reads, lists, diags, revs = [], [], [], []
    for i in range(0, 600, 4):
        qi, ti, dg = int(ntasks[i, 0]), int(ntasks[i, 1]), int(np.int16(np.uint16(ntasks[i, 2])))
        r = nqd[int(nqo[qi]):int(nqo[qi + 1])].copy()
        rev = (i // 4) % 3 == 2
        if rev:
            r = comp[r[::-1]]                      # the stored read is the reverse complement; the hit says "reverse strand"
        decoys = rng.integers(0, n_t, 3)
        reads.append(r)
        lists.append(np.array([ti] + list(decoys), np.uint32))
        diags.append(np.array([dg] + list(rng.integers(-100, 1500, 3)), np.int16))
        revs.append(np.array([1 if rev else 0, 0, 1 if rev else 0, 0], np.uint8))
    out["nucl_reads"] = np.concatenate(reads)
    out["nucl_read_off"] = np.cumsum([0] + [len(r) for r in reads]).astype(np.uint64)
    out["nucl_hit_targets"] = np.concatenate(lists)
    out["nucl_hit_diags"] = np.concatenate(diags)
    out["nucl_hit_rev"] = np.concatenate(revs)
    out["nucl_target_keys"] = nkeys
    for name, kw in NUCL_CONFIGS.items():
        texts, nal = [], []
        for k in range(len(reads)):
            txt, na, _ = ref.align_query_nucl(reads[k], 900 + k, ntd, nto, lists[k], nkeys[lists[k]], diags[k], revs[k], int(nto[-1]), **kw)
            texts.append(txt); nal.append(na)
        out["cfg_%s_text" % name] = np.array(texts)
        out["cfg_%s_naligned" % name] = np.array(nal, np.int64)
        print(name, "records:", sum(t.count(b"\n") for t in texts), "alignments:", sum(nal))
    # ---- DB triple: files written by the reference's DBWriter, and the `align` module's output DB for the default_a lists ----
    import tempfile
    tmp = tempfile.mkdtemp()
    out["aa2num_aa"], out["aa2num_nt"] = ref.aa2num()[:255], ref.aa2num(nucl=True)[:255]     # byte 255 is outside the reference's table
    wkeys = np.array([7, 3, 4000000000, 12, 5, 0], np.uint32)
    wents = [b"abc\n", b"", b"17\t250\t0\n8\t31\t-5\n", b"x" * 1000, b"\n", b"MKV\n"]
    ref.db_write(os.path.join(tmp, "w"), 7, wkeys, wents)
    out["dbw_keys"] = wkeys
    out["dbw_entries"] = np.array(wents, dtype=object).astype("S1000")
    out["dbw_entry_len"] = np.array([len(e) for e in wents])
    for suf, name in (("", "data"), (".index", "index"), (".dbtype", "dbtype")):
        out["dbw_file_" + name] = np.frombuffer(open(os.path.join(tmp, "w" + suf), "rb").read(), np.uint8)
    qkeys = 5000 + np.arange(len(qo) - 1)
    ref.db_write(os.path.join(tmp, "aln"), 5, qkeys, [bytes(t) for t in out["cfg_default_a_text"]])     # one thread: id order, as `align --threads 1`
    for suf, name in (("", "data"), (".index", "index"), (".dbtype", "dbtype")):
        out["alndb_file_" + name] = np.frombuffer(open(os.path.join(tmp, "aln" + suf), "rb").read(), np.uint8)
    # ---- `ungappedprefilter` output DB for the example queries (runFilterOnCpu, ungappedprefilter.cpp:418-478): the reference's
    #      ungapped_alignment scores, > 15, ordered by (score desc, key asc), at most 300 per query, written by its DBWriter
    pref_entries = []
    for qi in range(len(qo) - 1):
        q = qd[int(qo[qi]):int(qo[qi + 1])]
        sc = ref.ungapped(q, 1, td, to)
        ids = np.nonzero(sc > 15)[0]
        order = ids[np.lexsort((tkeys[ids], -sc[ids]))][:300]
        pref_entries.append(b"".join(b"%d\t%d\t0\n" % (int(tkeys[t]), int(sc[t])) for t in order))
    ref.db_write(os.path.join(tmp, "pref"), 7, qkeys, pref_entries)
    for suf, name in (("", "data"), (".index", "index"), (".dbtype", "dbtype")):
        out["prefdb_file_" + name] = np.frombuffer(open(os.path.join(tmp, "pref" + suf), "rb").read(), np.uint8)
    print("prefilter DB:", sum(e.count(b"\n") for e in pref_entries), "hits")
    np.savez_compressed(os.path.join(HERE, "align_v1.npz"), **out)
    print("wrote align_v1.npz", os.path.getsize(os.path.join(HERE, "align_v1.npz")))


if __name__ == "__main__":
    main()
