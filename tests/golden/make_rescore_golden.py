"""Regenerates tests/golden/rescore_v1.npz from the reference itself (oracle/_ref, built from /root/reference by oracle/Makefile):
the [123][123] ASCII substitution matrix rescorediagonal scores with (SubstitutionMatrix::createAsciiSubMat) and
DistanceCalculator::computeUngappedAlignment outputs for a set of ASCII sequence pairs, diagonals and all five rescore modes.
Run in the build container:  python tests/golden/make_rescore_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref  # noqa: E402

ref = Ref()
rng = np.random.default_rng(20260923)
m = ref.ascii_matrix()
letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXBZUacdklmn", np.uint8)
queries, targets, hits, expected = [], [], [], []
for rep in range(40):
    qL = int(rng.integers(1, 500))
    q = bytearray(rng.choice(letters, qL).tobytes())
    if rep % 5 == 0:
        q[0] = ord("*")
    queries.append(bytes(q))
    for k in range(6):
        tL = int(rng.integers(1, 500))
        t = bytearray(rng.choice(letters, tL).tobytes())
        if k % 2 == 0:
            n = int(rng.integers(1, min(qL, tL) + 1)); a = int(rng.integers(0, qL - n + 1)); b = int(rng.integers(0, tL - n + 1))
            t[b:b + n] = q[a:a + n]
            for j in rng.integers(0, n, n // 8):
                t[b + int(j)] = int(rng.choice(letters))
            diag = (a - b) & 0xffff
        else:
            diag = int(rng.integers(-tL, qL + 1)) & 0xffff
        if k == 3:
            t[-1] = ord("*")
        if k == 5:
            diag = int(rng.choice([5000, 60000, 0, 65535]))
        tid = len(targets)
        targets.append(bytes(t))
        hits.append((rep, tid, diag))
for mode in range(5):
    expected.append(np.stack([ref.rescore_diagonal(queries[qi], targets[ti], dg, mode) for qi, ti, dg in hits]))
qoff = np.zeros(len(queries) + 1, np.uint64); qoff[1:] = np.cumsum([len(x) for x in queries])
toff = np.zeros(len(targets) + 1, np.uint64); toff[1:] = np.cumsum([len(x) for x in targets])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rescore_v1.npz"), asciimat=m,
                    qdata=np.frombuffer(b"".join(queries), np.uint8), qoff=qoff, tdata=np.frombuffer(b"".join(targets), np.uint8), toff=toff,
                    hits=np.array(hits, np.int64), expected=np.stack(expected))
print("wrote rescore_v1.npz:", len(queries), "queries,", len(targets), "targets,", len(hits), "hits x 5 modes")
