#!/usr/bin/env python
"""tests/golden/make_paddeddb_golden.py -- fixtures for the padded-DB writer and the repeat masker (include/b200_db.h), made by the
reference itself:

  * a small FASTA (tandem repeats, low-complexity stretches, lower-case letters, odd letters, one- to four-residue sequences, equal
    lengths, the header styles Util::parseFastaHeader knows) -> `mmseqs createdb` -> `mmseqs makepaddedseqdb` with several parameter
    sets, every input and output file kept as bytes.  The binary is the unmodified AVX2 host built from /root/reference by
    integration/build_host.sh (integration/_build/mmseqs_avx2);
  * the likelihood-ratio matrix (ProbabilityMatrix of blosum62.out) and tantan::getProbabilities / Masker::maskSequence on a few
    sequences through oracle/_ref (lib/tantan/tantan.cpp and src/commons/Masker.cpp compiled in place, AVX2).

  python tests/golden/make_paddeddb_golden.py [path/to/mmseqs]
"""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

PARAM_SETS = {   # name -> (command-line flags, keyword arguments of mmseqs2_b200.db.make_padded_db)
    "default": ([], {}),
    "nomask": (["--mask", "0"], {"mask": 0}),
    "lower_runs": (["--mask-lower-case", "1", "--mask-n-repeat", "3"], {"mask_lower_case": 1, "mask_n_repeat": 3}),
    "prob05_nolookup": (["--mask-prob", "0.5", "--write-lookup", "0"], {"mask_prob": 0.5, "write_lookup": 0}),
}
SUFFIXES = ("", ".index", ".dbtype", "_h", "_h.index", "_h.dbtype", ".lookup", ".source")
AA = "ACDEFGHIKLMNPQRSTVWY"


def fasta(rng):
    def rnd(n):
        return "".join(AA[i] for i in rng.integers(0, 20, n))
    hdrs = ["sp|P12345|TEST_HUMAN some protein", "tr|Q9XYZ1|Q9XYZ1_MOUSE", "gi|12345|ref|NP_000001.1| hypothetical", "gnl|db|ident42 desc",
            "pat|US|123 x", "consensus_sp|P99999|CONS", "plainid description here", "ref|XP_1.1|", "pdb|1ABC|A", "gi|55", "pir||S12345 thing",
            "lcl|local1", "cl|c1|x", "consensus_plain", "bbs|777", "prf||0001 z", "gb|AAA1.1|locus", "UniRef90_A0A000 n=3", "id_with|bar|inside", "x"]
    recs = []
    for i in range(90):
        k = i % 9
        L = int(rng.integers(1, 400))
        if k == 0:
            s = rnd(L)
        elif k == 1:
            s = (rnd(int(rng.integers(1, 9))) * 200)[:L]                                         # tandem repeat
        elif k == 2:
            s = "A" * int(rng.integers(1, 30)) + rnd(L) + "A" * 12                                # opens with code-0 letters
        elif k == 3:
            s = rnd(L // 2) + "".join(rng.choice(list("QN"), L // 3 + 1)) + rnd(L // 3)           # low complexity
        elif k == 4:
            s = "".join(c.lower() if rng.random() < 0.3 else c for c in rnd(L))                    # soft-masked input
        elif k == 5:
            s = "".join(rng.choice(list(AA + "XBZUOJ*-")) for _ in range(L))                      # letters outside the alphabet
        elif k == 6:
            s = rnd(int(rng.integers(1, 5)))                                                      # tiny
        elif k == 7:
            s = rnd(120)                                                                          # equal lengths: the tie order
        else:
            s = rnd(L // 2) + "KKKKKKKKKK" + rnd(5) + "ggggggg" + rnd(L // 4)
        recs.append(">%s\n%s\n" % (hdrs[i % len(hdrs)] + ("_%d" % i if i >= len(hdrs) else ""), s))
    return "".join(recs)


def main():
    from oracle.pyoracle import Ref
    binary = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "integration", "_build", "mmseqs_avx2")
    ref = Ref()
    A = 21
    lr = np.zeros((A, A), np.float64)
    ref.lib.ref_tantan_matrix(lr.ctypes.data_as(ctypes.c_void_p))
    rng = np.random.default_rng(12)
    text = fasta(rng)
    out = {"lr": lr, "fasta": np.frombuffer(text.encode(), np.uint8)}
    with tempfile.TemporaryDirectory() as w:
        open(w + "/t.fasta", "w").write(text)
        subprocess.check_call([binary, "createdb", w + "/t.fasta", w + "/T", "-v", "1"], stdout=subprocess.DEVNULL)
        for suf in SUFFIXES:
            if os.path.exists(w + "/T" + suf):
                out["src" + suf] = np.frombuffer(open(w + "/T" + suf, "rb").read(), np.uint8)
        for name, (flags, _) in PARAM_SETS.items():
            subprocess.check_call([binary, "makepaddedseqdb", w + "/T", w + "/P_" + name, "-v", "1", "--threads", "3"] + flags, stdout=subprocess.DEVNULL)
            for suf in SUFFIXES:
                if os.path.exists(w + "/P_" + name + suf):
                    out["%s%s" % (name, suf)] = np.frombuffer(open(w + "/P_" + name + suf, "rb").read(), np.uint8)
    # masker on single sequences through oracle/_ref
    seqs = [s for s in text.split("\n") if s and not s.startswith(">")][:36]
    width = max(len(s) for s in seqs)
    codes = np.full((len(seqs), width), 255, np.uint8)
    probs = np.zeros((len(seqs), width), np.float32)
    masked = np.full((3, len(seqs), width), 255, np.uint8)
    counts = np.zeros((3, len(seqs)), np.int64)
    modes = [(1, 0.9, 0, 0), (1, 0.5, 1, 2), (0, 0.9, 1, 4)]          # (tantan, prob, lower case, n repeats)
    for i, s in enumerate(seqs):
        L = len(s)
        m = np.zeros(L, np.uint8)
        ref.lib.ref_mask_sequence(s.encode(), L, 0, ctypes.c_double(0.9), 0, 0, m.ctypes.data_as(ctypes.c_void_p))     # no masking: the codes
        codes[i, :L] = m
        p = np.zeros(L, np.float32)
        ref.lib.ref_tantan_probabilities(m.ctypes.data_as(ctypes.c_void_p), L, p.ctypes.data_as(ctypes.c_void_p))
        probs[i, :L] = p
        for k, (tt, pr, lc, nr) in enumerate(modes):
            o = np.zeros(L, np.uint8)
            counts[k, i] = ref.lib.ref_mask_sequence(s.encode(), L, tt, ctypes.c_double(pr), lc, nr, o.ctypes.data_as(ctypes.c_void_p))
            masked[k, i, :L] = o
    out.update(seq_text=np.array([s.encode() for s in seqs]), seq_codes=codes, seq_probs=probs, seq_masked=masked, seq_mask_counts=counts,
               seq_mask_modes=np.array(modes, np.float64))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "paddeddb_v1.npz"), **out)
    pad = out["default"]
    print("wrote paddeddb_v1.npz: %d sequences, padded data %d bytes, %d masked residues by default" % (text.count(">"), len(pad), int((pad >= 32).sum())))


if __name__ == "__main__":
    main()
