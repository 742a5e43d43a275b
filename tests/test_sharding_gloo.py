"""CPU, world_size 2 over gloo: the N>1 host logic (query decomposition + fixed-size hit-list gather) reproduces the
single-process result.  The per-rank "scan" is a deterministic stand-in; the GPU kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmseqs2_b200.sharding import decompose_by_residues, gather_hit_lists, merge_target_sharded

K = 7
HIT = np.dtype([("id", np.uint32), ("score", np.int32)])


def fake_scan(qids):
    hits = np.zeros((len(qids), K), HIT)
    n = np.zeros(len(qids), np.uint32)
    for i, q in enumerate(qids):
        m = 1 + (q * 5) % K
        n[i] = m
        hits[i, :m]["id"] = (np.arange(m) * 31 + q) % 1000
        hits[i, :m]["score"] = 200 - np.arange(m) - q
    return hits, n


def _worker(rank, world, port, qlens, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ranges = decompose_by_residues(qlens, world)
    s, e = ranges[rank]
    hits, n = fake_scan(list(range(s, e)))
    maxq = max(b - a for a, b in ranges)
    hb, cn = gather_hit_lists(hits, n, maxq, dist)
    if rank == 0:
        merged = []
        for r, (a, b) in enumerate(ranges):
            for i in range(b - a):
                merged.append(hb[r, i, :cn[r, i]].copy())
        ret.put([m.tolist() for m in merged])
    dist.barrier()
    dist.destroy_process_group()


def test_decompose_is_contiguous_balanced_and_complete():
    rng = np.random.default_rng(0)
    qlens = rng.integers(50, 2000, 1000)
    for world in (1, 2, 3, 8):
        r = decompose_by_residues(qlens, world)
        assert r[0][0] == 0 and r[-1][1] == len(qlens)
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        loads = [int(qlens[a:b].sum()) for a, b in r]
        assert max(loads) - min(loads) <= 2 * int(qlens.max())
    assert decompose_by_residues([], 2) == [(0, 0), (0, 0)]
    assert decompose_by_residues([5], 4)[-1][1] == 1


def test_two_rank_gather_equals_single_process():
    rng = np.random.default_rng(1)
    qlens = rng.integers(50, 2000, 37)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, qlens, ret)) for r in range(2)]
    for p in procs:
        p.start()
    merged = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hits, n = fake_scan(list(range(len(qlens))))
    assert len(merged) == len(qlens)
    for q in range(len(qlens)):
        exp = hits[q, :n[q]].view(np.int32).reshape(-1, 2).tolist()
        assert merged[q] == exp


# ---- target-sharded variant: each rank scans its slice of the DB, the top lists are merged ---------------------------------------
def _dense_scores(nq, n_targets):
    rng = np.random.default_rng(123)
    return rng.integers(0, 40, (nq, n_targets)).astype(np.int32)      # few distinct values: plenty of ties at the cut-off


def _topk(dense_row, ids, k, thr=15):
    sel = np.nonzero(dense_row > thr)[0]
    order = sel[np.lexsort((ids[sel], -dense_row[sel].astype(np.int64)))][:k]
    return ids[order], dense_row[order]


def _merge_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nq, nt, k = 6, 500, 20
    dense = _dense_scores(nq, nt)
    lo, hi = (0, 230) if rank == 0 else (230, nt)                       # uneven slices
    hits = np.zeros((nq, k), HIT)
    n = np.zeros(nq, np.uint32)
    for i in range(nq):
        ids, sc = _topk(dense[i, lo:hi], np.arange(hi - lo, dtype=np.uint32), k)
        n[i] = len(ids); hits["id"][i, :len(ids)] = ids; hits["score"][i, :len(ids)] = sc
    merged, nm = merge_target_sharded(hits, n, lo, k, dist)
    if rank == 1:
        ret.put((merged.tolist(), nm.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_target_sharded_merge_equals_unsharded_topk():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_merge_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    merged, nm = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nq, nt, k = 6, 500, 20
    dense = _dense_scores(nq, nt)
    for i in range(nq):
        ids, sc = _topk(dense[i], np.arange(nt, dtype=np.uint32), k)
        assert nm[i] == len(ids)
        assert [m[0] for m in merged[i][:nm[i]]] == ids.tolist() and [m[1] for m in merged[i][:nm[i]]] == sc.tolist()



# ---- pipelined gather (what bench.py --gpus N runs per step): several steps in flight, results in start order ------------------------
def _pipe_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmseqs2_b200.sharding import HitGather
    g = HitGather(5, K, dist, None)
    out = []
    for step in range(6):
        hits, n = fake_scan([100 * step + 10 * rank + i for i in range(3 + rank)])     # ranks hold different numbers of queries
        if len(g.started) >= 2:
            out.append(g.finish())
        g.start(hits, n)
    while g.started:
        out.append(g.finish())
    if rank == 0:
        ret.put([(hb.tolist(), cn.tolist()) for hb, cn in out])
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_gather_keeps_step_order_and_content():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_pipe_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    steps = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(steps) == 6
    for step, (hb, cn) in enumerate(steps):
        hb, cn = np.array(hb), np.array(cn)
        for rank in range(2):
            hits, n = fake_scan([100 * step + 10 * rank + i for i in range(3 + rank)])
            assert cn[rank, :3 + rank].tolist() == n.tolist() and (cn[rank, 3 + rank:] == 0).all()
            for i in range(3 + rank):
                assert hb[rank, i, :n[i]].tolist() == hits[i, :n[i]].view(np.int32).reshape(-1, 2).tolist()
