#!/usr/bin/env python
"""tests/golden/make_rescore_module_golden.py -- fixtures for the `rescorediagonal` module over DB files (include/b200_db.h), made by the
reference binary (the unmodified AVX2 host built from /root/reference by integration/build_host.sh):

  createdb (queries, targets: a few hundred sequences of the reference's example FASTAs that the k-mer prefilter connects) ->
  prefilter (-s 7.5, four threads: the result DB's data order differs from its key order) -> rescorediagonal --threads 1 with several
  parameter sets, plus one run of the target DB against itself (sameQTDB: identity hits).  Every input and output file is kept as bytes.

  python tests/golden/make_rescore_module_golden.py [path/to/mmseqs] [path/to/examples]
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PARAM_SETS = {   # name -> (command-line flags, keyword arguments of mmseqs2_b200.db.RescoreParams, same DB for queries and targets)
    "hamming": (["--rescore-mode", "0"], dict(rescore_mode=0), False),
    "hamming_minid": (["--rescore-mode", "0", "--min-seq-id", "0.3", "--sort-results", "1"], dict(rescore_mode=0, seq_id_thr=0.3, sort_results=1), False),
    "substitution": (["--rescore-mode", "1", "-e", "10", "--sort-results", "1"], dict(rescore_mode=1, eval_thr=10.0, sort_results=1), False),
    "alignment": (["--rescore-mode", "2", "-a", "--sort-results", "1", "-e", "10"], dict(rescore_mode=2, add_backtrace=1, sort_results=1, eval_thr=10.0), False),
    "alignment_cov": (["--rescore-mode", "2", "-c", "0.5", "--cov-mode", "1", "-e", "100", "--seq-id-mode", "1"],
                      dict(rescore_mode=2, cov_thr=0.5, cov_mode=1, eval_thr=100.0, seq_id_mode=1), False),
    "end_to_end": (["--rescore-mode", "3", "-e", "1000", "--min-aln-len", "30"], dict(rescore_mode=3, eval_thr=1000.0, aln_len_thr=30), False),
    "window_quality": (["--rescore-mode", "4", "-e", "10", "-a"], dict(rescore_mode=4, eval_thr=10.0, add_backtrace=1), False),
    "self_alignment": (["--rescore-mode", "2", "-e", "1e-5", "--sort-results", "1"], dict(rescore_mode=2, eval_thr=1e-5, sort_results=1), True),
}
DB_SUFFIXES = ("", ".index", ".dbtype")


def records(path):
    return open(path).read().split(">")[1:]


def read_data(path):
    """the data of a DB: <path>, or its split parts <path>.0, <path>.1, ... in order (FileUtil::findDatafiles)"""
    if os.path.exists(path):
        return open(path, "rb").read()
    parts, k = [], 0
    while os.path.exists("%s.%d" % (path, k)):
        parts.append(open("%s.%d" % (path, k), "rb").read())
        k += 1
    return b"".join(parts)


def main():
    binary = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "integration", "_build", "mmseqs_avx2")
    ex = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "integration", "_build", "examples")
    quiet = dict(stdout=subprocess.DEVNULL)
    out = {}
    with tempfile.TemporaryDirectory() as w:
        qrecs = records(os.path.join(ex, "QUERY.fasta"))[:24]
        trecs = records(os.path.join(ex, "DB.fasta"))[:4000]
        open(w + "/q.fasta", "w").write("".join(">" + r for r in qrecs))
        open(w + "/t0.fasta", "w").write("".join(">" + r for r in trecs))
        subprocess.check_call([binary, "createdb", w + "/q.fasta", w + "/Q", "-v", "1"], **quiet)
        subprocess.check_call([binary, "createdb", w + "/t0.fasta", w + "/T0", "-v", "1"], **quiet)
        subprocess.check_call([binary, "prefilter", w + "/Q", w + "/T0", w + "/pref0", "-v", "1", "--threads", "4", "-s", "7.5"], **quiet)
        # keep the targets the prefilter connects to a query (best 30 per query) + a few unconnected ones
        keep = set(range(0, 4000, 97))
        data = read_data(w + "/pref0")
        for line in open(w + "/pref0.index"):
            _, off, ln = (int(x) for x in line.split())
            rows = [r.split(b"\t") for r in data[off:off + ln - 1].splitlines()]
            rows.sort(key=lambda r: -int(r[1]))
            keep.update(int(r[0]) for r in rows[:30])
        open(w + "/t.fasta", "w").write("".join(">" + trecs[k] for k in sorted(keep)))
        subprocess.check_call([binary, "createdb", w + "/t.fasta", w + "/T", "-v", "1"], **quiet)
        subprocess.check_call([binary, "prefilter", w + "/Q", w + "/T", w + "/pref", "-v", "1", "--threads", "4", "-s", "7.5"], **quiet)
        subprocess.check_call([binary, "prefilter", w + "/T", w + "/T", w + "/pref_tt", "-v", "1", "--threads", "4", "-s", "5", "--max-seqs", "20"], **quiet)
        for name in ("Q", "T", "pref", "pref_tt"):
            out[name] = np.frombuffer(read_data(w + "/" + name), np.uint8)       # a split result DB (.0 .1 ..) is kept as the one file it reads as
            for suf in DB_SUFFIXES[1:]:
                out[name + suf] = np.frombuffer(open(w + "/" + name + suf, "rb").read(), np.uint8)
        for name, (flags, _, same) in PARAM_SETS.items():
            q, p = ("T", "pref_tt") if same else ("Q", "pref")
            subprocess.check_call([binary, "rescorediagonal", w + "/" + q, w + "/T", w + "/" + p, w + "/R_" + name, "-v", "1", "--threads", "1"] + flags, **quiet)
            for suf in DB_SUFFIXES:
                out["out_" + name + suf] = np.frombuffer(open(w + "/R_" + name + suf, "rb").read(), np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rescore_module_v1.npz"), **out)
    print("wrote rescore_module_v1.npz: %d queries, %d targets; output bytes: %s" % (
        len(qrecs), len(keep), {n: len(out["out_" + n]) for n in PARAM_SETS}))


if __name__ == "__main__":
    main()
