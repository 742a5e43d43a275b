"""tests/golden/make_golden.py -- regenerate the committed fixtures from the REFERENCE's own code.

Runs only in the build container (needs oracle/_ref/libmmseqs_ref.so, i.e. /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/blosum62.npz (integer matrix at bitFactor 2, background, alphabet) and
tests/golden/hotpath_v1.npz (seeded inputs + the reference outputs for A1, A2, A3-A5).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref, pack_targets  # noqa: E402
from mmseqs2_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def read_fasta(path, limit):
    seqs, cur = [], []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if cur:
                    seqs.append("".join(cur)); cur = []
                    if len(seqs) >= limit:
                        break
            else:
                cur.append(line.strip())
    if cur and len(seqs) < limit:
        seqs.append("".join(cur))
    return seqs


def make_examples(ref):
    """tests/golden/examples_v1.npz: real proteins (BASELINE config[0] data: the reference's examples/QUERY.fasta and
    DB.fasta, first 40 queries x first 600 targets, residues mapped with the reference's aa2num) and the reference's outputs"""
    a2n = ref.aa2num()
    qs = [a2n[np.frombuffer(s.encode(), np.uint8)] for s in read_fasta("/root/reference/examples/QUERY.fasta", 40)]
    pool = [a2n[np.frombuffer(s.encode(), np.uint8)] for s in read_fasta("/root/reference/examples/DB.fasta", 8000)]
    pd_, po_ = pack_targets(pool)
    chosen = []
    for q in qs:                                  # make sure real homologs are in the 600-target mini DB
        sc = ref.ungapped(q, 1, pd_, po_)
        top = np.argsort(-sc, kind="stable")[:12]
        chosen.extend(int(t) for t in top if sc[t] > 15)
    chosen = sorted(set(chosen))
    rest = [i for i in range(len(pool)) if i not in set(chosen)]
    ids600 = sorted(chosen + rest[:max(0, 600 - len(chosen))])[:600] if len(chosen) < 600 else chosen[:600]
    ts = [pool[i] for i in ids600]
    td, to = pack_targets(ts)
    qd, qo = pack_targets(qs)
    ung = np.stack([ref.ungapped(q, 1, td, to) for q in qs])
    aln = np.stack([ref.ssw_align(q, 1, td, to, mode=1)[0][:, :6] for q in qs])
    # end-to-end hit lists the way `search --prefilter-mode 1` produces them on this mini DB:
    #   ungappedprefilter: score > 15, order (score desc, id asc), top 300          (ungappedprefilter.cpp:450-478)
    #   align: ssw_align mode 1, accept evalue <= 1e-3, order Matcher::compareHits  (Alignment.cpp:346-405, Matcher.h:161-172)
    db_res = int(to[-1])
    tlen = np.diff(to)
    hit_off, hit_ids, hit_rows, min_score = [0], [], [], []
    for qi, q in enumerate(qs):
        sc = ung[qi].astype(np.int64)
        ids = np.nonzero(sc > 15)[0]
        ids = ids[np.lexsort((ids, -sc[ids]))][:300]
        sd, so = pack_targets([ts[i] for i in ids]) if len(ids) else (np.zeros(0, np.uint8), np.zeros(1, np.int64))
        accepted = []
        if len(ids):
            a, ev, _ = ref.ssw_align(q, 1, sd, so, mode=1, eval_thr=1e-3, db_residues=db_res)
            for k, t in enumerate(ids):
                if a[k, 4] != -1 and ev[k] <= 1e-3:
                    accepted.append((ev[k], -int(a[k, 0]), int(tlen[t]), int(t), a[k, :6].copy()))
        accepted.sort(key=lambda r: r[:4])
        hit_ids.extend(r[3] for r in accepted)
        hit_rows.extend(r[4] for r in accepted)
        hit_off.append(len(hit_ids))
        ms = 1
        while ref.evalue(11, 1, db_res, ms, len(q)) > 1e-3:
            ms += 1
        min_score.append(ms)
    np.savez_compressed(os.path.join(HERE, "examples_v1.npz"), qdata=qd, qoff=qo, tdata=td, toff=to, ungapped=ung.astype(np.uint8),
                        align=aln.astype(np.int32), hit_off=np.array(hit_off, np.int64), hit_ids=np.array(hit_ids, np.int64),
                        hit_rows=np.array(hit_rows, np.int32).reshape(-1, 6), min_score=np.array(min_score, np.int32))
    print("search fixture:", len(hit_ids), "accepted hits, min scores", min(min_score), "-", max(min_score))
    print("examples fixture:", len(qs), "x", len(ts), "max ungapped", int(ung.max()), "word pairs", int(aln[:, :, 5].sum()))


def make_nucl(ref):
    """tests/golden/nucl_v1.npz: reads against genome pieces, outputs of the reference's BandedNucleotideAligner::align"""
    rng = np.random.default_rng(5150)
    targets = [synth.nucl_genome(rng, int(n)) for n in rng.integers(200, 3000, 40)]
    targets[3][::17] = 4                       # X residues
    reads, tasks = synth.nucl_reads(rng, targets, 600, 150, subst=0.04, indel=0.01)
    for i in range(0, 600, 9):                 # off-diagonal seeds, short reads, perfect full-length copies
        tasks[i, 2] = (int(tasks[i, 2]) + int(rng.integers(-6, 7))) & 0xffff
    for i in range(0, 600, 50):
        t = int(tasks[i, 1]); reads[i] = targets[t].copy(); tasks[i, 2] = 0
    for i in range(5, 600, 40):
        reads[i] = reads[i][:int(rng.integers(5, 40))]
    td, to = pack_targets(targets)
    outs, bts, cigs, coff = [], [], [], [0]
    for (qi, ti, dg) in tasks:
        d = int(dg) if dg < 32768 else int(dg) - 65536
        o, cg, bt = ref.nucl_align(reads[qi], targets[ti], d)
        outs.append(o); bts.append(bt); cigs.append(cg); coff.append(coff[-1] + len(cg))
    qd, qo = pack_targets(reads)
    np.savez_compressed(os.path.join(HERE, "nucl_v1.npz"), tdata=td, toff=to, qdata=qd, qoff=qo, tasks=tasks, out=np.array(outs, np.int32),
                        cigars=np.concatenate(cigs).astype(np.uint32), cigar_off=np.array(coff, np.int64),
                        bt=np.array(bts))
    print("nucl fixture:", len(tasks), "alignments,", int(sum(o[6] > 1 for o in outs)), "with gaps")


def main():
    ref = Ref()
    mat, pb, n2a = ref.matrix()
    nmat, npb, nn2a = ref.matrix(nucl=True)
    np.savez_compressed(os.path.join(HERE, "blosum62.npz"), mat=mat, pback=pb, alphabet=np.frombuffer(n2a.encode(), np.uint8),
                        nucl_mat=nmat, nucl_pback=npb, nucl_alphabet=np.frombuffer(nn2a.encode(), np.uint8))
    rng = np.random.default_rng(20260922)
    bg = synth.background(pb)
    out = {}
    qlens = [1, 7, 33, 64, 65, 150, 255, 256, 257, 350, 511, 700, 1030]
    queries = [synth.random_seqs(rng, 1, bg, mean=L, sigma=0, lo=L, hi=L, normal=True)[0] for L in qlens]
    # X residues in a few queries
    queries[3][::7] = 20
    res, off = synth.random_seqs(rng, 400, bg, mean=250, sigma=0.7, lo=1, hi=1500)
    synth.plant_homologs(rng, res, off, [q for q in queries if len(q) > 30], bg, frac=0.5, subst=0.25, indel=0.04)
    # a near-identical long pair to force word mode, and targets starting with X
    tg = synth.split(res.copy(), off)
    tg[0] = queries[-1].copy()
    tg[1] = synth.mutate(rng, queries[-2], bg, 0.1, 0.02)
    tg[2] = np.concatenate([np.full(3, 20, np.uint8), tg[2]])
    td, to = pack_targets(tg)
    out["tdata"], out["toff"] = td, to
    out["nq"] = np.array(len(queries))
    for qi, q in enumerate(queries):
        out["q%d" % qi] = q
        for cbf in (0, 1):
            key = "q%d_cb%d_" % (qi, cbf)
            out[key + "ungapped"] = ref.ungapped(q, cbf, td, to)
            out[key + "endpos"] = ref.sw_score_endpos(q, cbf, td, to)
            aln, ev, _ = ref.ssw_align(q, cbf, td, to, mode=1)
            out[key + "align"] = aln[:, :6]
            if cbf == 1:   # alignment mode 2: backtrace string + identity count (banded_sw + computerBacktrace)
                aln2, _, bts = ref.ssw_align(q, cbf, td, to, mode=2, want_bt=True)
                out[key + "bt"] = np.array(bts)
                out[key + "ident"] = aln2[:, 6]
        # per-diagonal scorer
        nh = 1500
        ids = rng.integers(0, len(tg), nh).astype(np.uint32)
        dg = rng.integers(-len(q) - 5, 600, nh).astype(np.int16).view(np.uint16)
        ids[:6] = [0, 1, 0, 1, 0, 1]          # the planted (near-)identical targets: exercises the 255 clamp (T2)
        dg[:6] = np.array([0, 0, 1, -1, -2, 3], np.int16).view(np.uint16)
        f = ref.comp_bias(q)
        out["q%d_diag_ids" % qi], out["q%d_diag_dg" % qi] = ids, dg
        for cbf in (0, 1):
            c, r = ref.diag(q, f if cbf else None, td, to, ids, dg)
            out["q%d_cb%d_diag_counts" % (qi, cbf)] = c
            out["q%d_cb%d_diag_raw" % (qi, cbf)] = r
        out["q%d_compbias" % qi] = f
    np.savez_compressed(os.path.join(HERE, "hotpath_v1.npz"), **out)
    make_nucl(ref)
    make_examples(ref)
    print("wrote fixtures:", {k: os.path.getsize(os.path.join(HERE, k)) for k in ("blosum62.npz", "hotpath_v1.npz")})


if __name__ == "__main__":
    main()
