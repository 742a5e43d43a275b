"""tests/golden/make_profile_golden.py -- profile (PSSM) query fixtures from the REFERENCE's own HMM_PROFILE code paths
(ssw_init / ssw_align / ungapped_alignment with a profile query, UngappedAlignment::createProfile + align), via oracle/_ref.
Build container only:  python tests/golden/make_profile_golden.py  ->  tests/golden/profile_v1.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref, pack_targets  # noqa: E402
from mmseqs2_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make_pssm(rng, mat, q, sharp, lo=-128):
    """a position-specific table around a seed sequence: scaled matrix column + noise, a few strongly conserved and a few
    near-flat positions; int8 [20][L]"""
    L = len(q)
    base = mat[:20, :][:, q].astype(np.float64)
    p = np.rint(base * sharp + rng.normal(0, 1.2, base.shape))
    cons = rng.random(L) < 0.15
    p[:, cons] = np.rint(base[:, cons] * 2.5) - 3
    flat = rng.random(L) < 0.1
    p[:, flat] = rng.integers(-1, 2, (20, int(flat.sum())))
    return np.clip(p, lo, 127).astype(np.int8)


def main():
    ref = Ref()
    mat, pb, _ = ref.matrix()
    ex = np.load(os.path.join(HERE, "examples_v1.npz"))
    rng = np.random.default_rng(424242)
    bg = synth.background(pb)
    seqs = [ex["tdata"][int(ex["toff"][i]):int(ex["toff"][i + 1])] for i in range(300)]
    qsrc = [ex["qdata"][int(ex["qoff"][i]):int(ex["qoff"][i + 1])] for i in (0, 3, 7, 11, 19)]
    qsrc.append(synth.random_seqs(rng, 1, bg, mean=700, sigma=0, lo=700, hi=700, normal=True)[0])     # two query tiles
    qsrc.append(synth.random_seqs(rng, 1, bg, mean=5, sigma=0, lo=5, hi=5, normal=True)[0])
    qsrc.append(synth.random_seqs(rng, 1, bg, mean=129, sigma=0, lo=129, hi=129, normal=True)[0])
    sharps = [1.0, 0.6, 1.4, 1.0, 2.0, 1.0, 1.0, 3.0]
    pssms = [make_pssm(rng, mat, q, s) for q, s in zip(qsrc, sharps)]
    pssms[1][:, ::9] = -60                     # large profile bias (byte mode rarely usable)
    pssms[7] = np.clip(pssms[7].astype(np.int32) * 2, -128, 127).astype(np.int8)   # scores far beyond a matrix: word mode
    cons = [p.argmax(0).astype(np.uint8) for p in pssms]
    # targets: real proteins + mutated copies of the consensus sequences (strong hits, some with X)
    for k, c in enumerate(cons):
        if len(c) > 20:
            seqs[2 * k] = synth.mutate(rng, c, bg, 0.15, 0.02)
            seqs[2 * k + 1] = c.copy()
    seqs[40][::11] = 20
    td, to = pack_targets(seqs)
    out = {"tdata": td, "toff": to, "n": np.array(len(pssms))}
    for k, (p, c) in enumerate(zip(pssms, cons)):
        out["pssm%d" % k], out["cons%d" % k] = p, c
        out["ungapped%d" % k] = ref.profile_align(p, c, td, to, mode=-1)[0][:, 0]
        a1, _, _ = ref.profile_align(p, c, td, to, mode=1)
        out["align%d" % k] = a1[:, :6]
        a2, _, bts = ref.profile_align(p, c, td, to, mode=2, want_bt=True)
        out["bt%d" % k] = np.array(bts)
        out["ident%d" % k] = a2[:, 6]
        nh = 800
        ids = rng.integers(0, len(seqs), nh).astype(np.uint32)
        dg = rng.integers(-len(c) - 3, 500, nh).astype(np.int16).view(np.uint16)
        ids[:4] = [2 * k % len(seqs), (2 * k + 1) % len(seqs), 2 * k % len(seqs), (2 * k + 1) % len(seqs)]
        dg[:4] = np.array([0, 0, 1, -1], np.int16).view(np.uint16)
        cnt, raw = ref.profile_diag(p, c, td, to, ids, dg)
        out["diag_ids%d" % k], out["diag_dg%d" % k], out["diag_counts%d" % k], out["diag_raw%d" % k] = ids, dg, cnt, raw
        print("pssm", k, "L", len(c), "bias", int(-p.min()), "max ungapped", int(out["ungapped%d" % k].max()), "word pairs",
              int(a1[:, 5].sum()), "clamped diag", int((cnt == 255).sum()))
    np.savez_compressed(os.path.join(HERE, "profile_v1.npz"), **out)
    print("wrote profile_v1.npz", os.path.getsize(os.path.join(HERE, "profile_v1.npz")))


if __name__ == "__main__":
    main()
