"""Query sharding across GPUs and the top-k hit gather.

The path shards by query: every (query, target) pair is independent and the DB is replicated, so the only exchange is
the gather of finished, fixed-size per-query hit lists.  decompose_by_residues mirrors the reference's work split
(DBReader::decomposeDomainByAminoAcid, src/commons/DBReader.cpp:1108: contiguous ranges balanced by residue count).
"""
import numpy as np


def decompose_by_residues(qlens, world):
    """-> list of (start, end) contiguous query ranges, one per rank, balanced by sum of lengths"""
    qlens = np.asarray(qlens, np.int64)
    n = len(qlens)
    total = int(qlens.sum())
    bounds = [0]
    csum = np.cumsum(qlens)
    for r in range(1, world):
        target = total * r / world
        idx = int(np.searchsorted(csum, target, side="left")) + 1 if n else 0
        idx = max(bounds[-1], min(n, idx))
        bounds.append(idx)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def gather_hit_lists(hits, n_hits, max_queries_per_rank, dist=None, device=None):
    """hits: structured (id u32, score i32) array [nq_local][k]; n_hits: [nq_local].
    Returns (hits_all [world][max_q][k], n_all [world][max_q]) on every rank; one all_gather of fixed-size records."""
    import torch
    k = hits.shape[1] if hits.ndim == 2 else 0
    buf = np.zeros((max_queries_per_rank, k, 2), np.int32)
    cnt = np.zeros(max_queries_per_rank, np.int32)
    nq = hits.shape[0]
    if nq:
        buf[:nq] = hits.view(np.int32).reshape(nq, k, 2)
        cnt[:nq] = n_hits
    if dist is None or not dist.is_initialized():
        return buf[None], cnt[None]
    world = dist.get_world_size()
    t = torch.from_numpy(np.concatenate([buf.reshape(-1), cnt]))
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    flat = torch.stack(out).cpu().numpy()
    hb = flat[:, :buf.size].reshape(world, max_queries_per_rank, k, 2)
    cn = flat[:, buf.size:]
    return hb, cn
