"""Synthetic protein data of the BASELINE shapes: i.i.d. residues from the BLOSUM62 background,
log-normal or normal lengths, planted mutated homologs.  numpy only; seeds are explicit so every rank / test / bench
run regenerates identical bytes."""
import numpy as np


def background(pback):
    p = np.array(pback[:20], np.float64)  # X excluded (prob 1e-5 in the matrix file -> 0 here)
    return p / p.sum()


def random_seqs(rng, n, bg, mean=300.0, sigma=0.6, lo=30, hi=5000, normal=False):
    """-> (residues uint8 concatenated, offsets uint64[n+1])"""
    if normal:
        lens = np.clip(np.rint(rng.normal(mean, sigma, n)), lo, hi).astype(np.int64)
    else:
        lens = np.clip(np.rint(rng.lognormal(np.log(mean), sigma, n)), lo, hi).astype(np.int64)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    cdf = np.cumsum(bg)
    res = np.searchsorted(cdf, rng.random(total), side="right").astype(np.uint8)
    np.minimum(res, 19, out=res)
    return res, off


def mutate(rng, seq, bg, subst=0.3, indel=0.02):
    """copy of seq with substitutions and short indels (vectorised; deterministic for a given rng state)"""
    seq = np.asarray(seq, np.uint8)
    n = len(seq)
    cdf = np.cumsum(bg)
    out = seq.copy()
    sub = rng.random(n) < subst
    out[sub] = np.minimum(np.searchsorted(cdf, rng.random(int(sub.sum())), side="right"), 19).astype(np.uint8)
    u = rng.random(n)
    keep = u >= indel / 2                      # deletions
    ins_at = np.nonzero((u >= indel / 2) & (u < indel))[0]   # insertions before these positions
    pieces = []
    last = 0
    for pos in ins_at:
        pieces.append(out[last:pos][keep[last:pos]])
        k = int(rng.integers(1, 6))
        pieces.append(np.minimum(np.searchsorted(cdf, rng.random(k), side="right"), 19).astype(np.uint8))
        last = pos
    pieces.append(out[last:][keep[last:]])
    res = np.concatenate(pieces) if pieces else out
    if len(res) == 0:
        res = seq[:1].copy()
    return res.astype(np.uint8)


def mutate_many(rng, seq, n, bg, ident_lo=0.2, ident_hi=0.9, indel=0.02):
    """n mutated copies of seq at identities uniform in [ident_lo, ident_hi], fully vectorised: substitutions by a mask over an
    [n][L] tile, deletions (indel/2 per residue) and 3-residue insertions (indel/6 per residue) by repeating every cell 0, 1 or 4
    times.  -> (residues concatenated, lengths[n]).  For the 1 M-pair SW workload."""
    seq = np.asarray(seq, np.uint8)
    L = len(seq)
    cdf = np.cumsum(bg)
    rows = np.tile(seq, (n, 1))
    subst = 1.0 - rng.uniform(ident_lo, ident_hi, n)
    mask = rng.random((n, L)) < subst[:, None]
    rows[mask] = np.minimum(np.searchsorted(cdf, rng.random(int(mask.sum())), side="right"), 19).astype(np.uint8)
    u = rng.random((n, L))
    counts = np.ones((n, L), np.int64)
    counts[u < indel / 2] = 0
    counts[(u >= indel / 2) & (u < indel / 2 + indel / 6)] = 4
    counts[:, 0] = np.maximum(counts[:, 0], 1)                 # never an empty sequence
    flat_counts = counts.reshape(-1)
    data = np.repeat(rows.reshape(-1), flat_counts)
    # the 2nd..4th copy of a repeated cell is an inserted random residue
    starts = np.cumsum(flat_counts) - flat_counts
    within = np.arange(len(data)) - np.repeat(starts, flat_counts)
    ins = within > 0
    data[ins] = np.minimum(np.searchsorted(cdf, rng.random(int(ins.sum())), side="right"), 19).astype(np.uint8)
    return data, counts.sum(1)


def mutate_many_torch(gen, seq, n, ident_lo=0.2, ident_hi=0.9, indel=0.02, device="cuda"):
    """mutate_many on the GPU (synthetic-data plumbing for the 1 M-pair workload: 5e8 residues of homologs take minutes in numpy);
    substituted / inserted residues are uniform over the 20 letters.  gen: torch.Generator on `device`.  -> (uint8 tensor, lengths)"""
    import torch
    q = torch.as_tensor(np.asarray(seq, np.uint8), device=device)
    L = q.numel()
    rows = q.repeat(n, 1)
    subst = 1.0 - (ident_lo + (ident_hi - ident_lo) * torch.rand(n, device=device, generator=gen))
    mask = torch.rand((n, L), device=device, generator=gen) < subst[:, None]
    rnd = torch.randint(0, 20, (n, L), device=device, generator=gen, dtype=torch.uint8)
    rows = torch.where(mask, rnd, rows)
    u = torch.rand((n, L), device=device, generator=gen)
    counts = torch.ones((n, L), dtype=torch.int64, device=device)
    counts[u < indel / 2] = 0
    counts[(u >= indel / 2) & (u < indel / 2 + indel / 6)] = 4
    counts[:, 0].clamp_(min=1)
    flat = counts.reshape(-1)
    data = torch.repeat_interleave(rows.reshape(-1), flat)
    starts = torch.cumsum(flat, 0) - flat
    within = torch.arange(data.numel(), device=device) - torch.repeat_interleave(starts, flat)
    ins = within > 0
    data = torch.where(ins, torch.randint(0, 20, (data.numel(),), device=device, generator=gen, dtype=torch.uint8), data)
    return data, counts.sum(1)


def plant_homologs(rng, res, off, queries, bg, frac=0.01, subst=0.3, indel=0.02):
    """overwrite a window of ~frac of the targets with a mutated copy of a random query segment (in place)"""
    n = len(off) - 1
    k = max(1, int(n * frac))
    ids = rng.choice(n, size=k, replace=False)
    for t in ids:
        q = queries[int(rng.integers(0, len(queries)))]
        a = int(rng.integers(0, max(1, len(q) - 20)))
        b = int(rng.integers(a + 10, len(q) + 1)) if len(q) > a + 10 else len(q)
        m = mutate(rng, q[a:b], bg, subst, indel)
        tl = int(off[t + 1] - off[t])
        m = m[:tl]
        s = int(rng.integers(0, tl - len(m) + 1))
        res[int(off[t]) + s:int(off[t]) + s + len(m)] = m
    return ids


def split(res, off):
    return [res[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]


def pack(seqs):
    """list of uint8 arrays -> (concatenated residues, uint64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, np.uint64)
    if len(seqs):
        off[1:] = np.cumsum([len(x) for x in seqs])
    data = np.concatenate(seqs).astype(np.uint8) if len(seqs) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(data), off


def nucl_genome(rng, n):
    """i.i.d. A,C,T,G (codes 0..3)"""
    return rng.integers(0, 4, n).astype(np.uint8)


def nucl_mutate(rng, seq, subst=0.02, indel=0.002):
    seq = np.asarray(seq, np.uint8)
    out = seq.copy()
    sub = rng.random(len(seq)) < subst
    out[sub] = rng.integers(0, 4, int(sub.sum())).astype(np.uint8)
    u = rng.random(len(seq))
    keep = u >= indel / 2
    pieces, last = [], 0
    for pos in np.nonzero((u >= indel / 2) & (u < indel))[0]:
        pieces.append(out[last:pos][keep[last:pos]])
        pieces.append(rng.integers(0, 4, int(rng.integers(1, 4))).astype(np.uint8))
        last = pos
    pieces.append(out[last:][keep[last:]])
    res = np.concatenate(pieces)
    return res if len(res) else seq[:1].copy()


def nucl_reads(rng, targets, n_reads, read_len=150, subst=0.02, indel=0.002):
    """reads sampled from the targets; -> (reads, tasks[(read, target, diagonal_u16)]) with the true diagonal
    (diagonal = query position - target position as the prefilter stores it, UngappedAlignment.cpp:423-437)"""
    reads, tasks = [], []
    for i in range(n_reads):
        t = int(rng.integers(0, len(targets)))
        tl = len(targets[t])
        L = min(read_len, tl)
        pos = int(rng.integers(0, tl - L + 1))
        reads.append(nucl_mutate(rng, targets[t][pos:pos + L], subst, indel))
        tasks.append((i, t, (-pos) & 0xffff))
    return reads, np.array(tasks, np.int64)
