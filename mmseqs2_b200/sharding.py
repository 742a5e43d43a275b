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


class HitGather:
    """The same exchange, pipelined: start(step s) enqueues pinned-host -> device, one all_gather_into_tensor (NCCL over NVLink) and the
    copy back on a high-priority stream of its own and returns at once; finish() hands out the gathered lists of the oldest started
    step.  The caller keeps two scans in flight on the library's stream meanwhile, so the collective (a few hundred KB) and its
    host-side staging never sit between two scans.  Fixed-size records: [max_q][k][2] int32 + [max_q] counts per rank."""

    def __init__(self, max_queries_per_rank, k, dist, device, depth=3):
        import torch
        self.torch, self.dist, self.device = torch, dist, device
        self.mq, self.k = max_queries_per_rank, k
        self.world = dist.get_world_size()
        self.n = self.mq * self.k * 2 + self.mq
        self.cuda = device is not None and torch.device(device).type == "cuda"     # CPU (gloo) path: same protocol, for the tests
        self.stream = torch.cuda.Stream(device=device, priority=-1) if self.cuda else None
        self.slots = []
        for _ in range(depth):
            h_in, h_out = torch.zeros(self.n, dtype=torch.int32), torch.zeros(self.world * self.n, dtype=torch.int32)
            self.slots.append({"h_in": h_in.pin_memory() if self.cuda else h_in,
                               "d_in": torch.zeros(self.n, dtype=torch.int32, device=device) if self.cuda else None,
                               "d_out": torch.zeros(self.world * self.n, dtype=torch.int32, device=device) if self.cuda else None,
                               "h_out": h_out.pin_memory() if self.cuda else h_out,
                               "ev": torch.cuda.Event() if self.cuda else None, "work": None})
        self.started, self.head = [], 0

    def start(self, hits, n_hits):
        torch = self.torch
        sl = self.slots[self.head % len(self.slots)]
        self.head += 1
        buf = sl["h_in"].numpy()
        buf[:] = 0
        nq = hits.shape[0]
        if nq:
            buf[:self.mq * self.k * 2].reshape(self.mq, self.k, 2)[:nq] = hits.view(np.int32).reshape(nq, self.k, 2)
            buf[self.mq * self.k * 2:][:nq] = n_hits
        if self.cuda:
            with torch.cuda.stream(self.stream):
                sl["d_in"].copy_(sl["h_in"], non_blocking=True)
                self.dist.all_gather_into_tensor(sl["d_out"], sl["d_in"])
                sl["h_out"].copy_(sl["d_out"], non_blocking=True)
                sl["ev"].record(self.stream)
        else:
            outs = list(sl["h_out"].view(self.world, self.n).unbind(0))
            sl["work"] = self.dist.all_gather(outs, sl["h_in"], async_op=True)
        self.started.append(sl)

    def finish(self):
        """-> (hits_all [world][max_q][k][2], n_all [world][max_q]) of the oldest started step"""
        sl = self.started.pop(0)
        if self.cuda:
            sl["ev"].synchronize()
        else:
            sl["work"].wait()
        flat = sl["h_out"].numpy().reshape(self.world, self.n).copy()     # the slot is reused by a later start()
        return flat[:, :self.mq * self.k * 2].reshape(self.world, self.mq, self.k, 2), flat[:, self.mq * self.k * 2:]

    def drain(self):
        out = None
        while self.started:
            out = self.finish()
        return out


def merge_target_sharded(hits, n_hits, id_offset, k, dist=None, device=None):
    """The other split of SURVEY 8e, for a DB too large for one GPU: every rank holds a slice of the TARGETS (its local ids start
    at id_offset in the global numbering) and has scanned all queries against it.  hits [nq][k_local] / n_hits [nq] are the local
    top lists; the result on every rank is the global top-k per query, ordered like hit_t::compareHitsByScoreAndId (score desc,
    global id asc) -- the deterministic tie-break that makes the merged list independent of the number of ranks.
    One all_gather of fixed-size records, then a k-way merge (here: one sort of world * k_local records per query)."""
    nq = hits.shape[0]
    local = hits.copy()
    for i in range(nq):
        local["id"][i, :int(n_hits[i])] += np.uint32(id_offset)
    hb, cn = gather_hit_lists(local, n_hits, nq, dist, device)
    world = hb.shape[0]
    out = np.zeros((nq, k), hits.dtype)
    n_out = np.zeros(nq, np.uint32)
    for i in range(nq):
        ids = np.concatenate([hb[r, i, :int(cn[r, i]), 0] for r in range(world)]).view(np.uint32) if world else np.zeros(0, np.uint32)
        sc = np.concatenate([hb[r, i, :int(cn[r, i]), 1] for r in range(world)]) if world else np.zeros(0, np.int32)
        order = np.lexsort((ids, -sc.astype(np.int64)))[:k]
        n_out[i] = len(order)
        out["id"][i, :len(order)] = ids[order]
        out["score"][i, :len(order)] = sc[order]
    return out, n_out

