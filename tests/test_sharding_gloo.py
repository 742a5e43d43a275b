"""CPU, world_size 2 over gloo: the N>1 host logic (query decomposition + fixed-size hit-list gather) reproduces the
single-process result.  The per-rank "scan" is a deterministic stand-in; the GPU kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmseqs2_b200.sharding import decompose_by_residues, gather_hit_lists

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
