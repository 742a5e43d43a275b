#!/usr/bin/env python
"""scripts/multi_gpu_configs.py -- BASELINE configs[3] and [4] on N GPUs of one node (one process per GPU, torchrun).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_configs.py

config[3]  full easy-search shape: queries vs a 20 M-sequence UniRef-shaped synthetic DB (log-normal lengths, ~7 GB of residues),
           DB replicated on every GPU (fully resident), queries sharded by rank -- on a query SUBSAMPLE (--q3 per rank; the scan is
           linear in queries): ungapped prefilter of every query against the whole DB, then the `align` step (-a, -e 1e-3) on the hit
           lists; hit lists gathered with one NCCL all_gather per batch (pipelined as in bench.py).
config[4]  nucleotide mode: 150-bp reads (2 % substitutions, 0.2 % indels, both strands) sampled from a 5 Gbp synthetic genome cut
           into 32 000-bp targets (the aligner's limit is 32 767; `splitsequence` cuts longer ones upstream), genome replicated on every GPU, reads sharded by rank (1 M reads / N per rank), gapped
           nucleotide aligner on the prefilter diagonal of every read (nucleotide.out, gap 5/2, zdrop 40, band 64).

Synthetic data is generated on the GPU with torch (data plumbing: 7e9 residues take minutes in numpy) and handed to the library as
host buffers, like every other caller.  Rank 0 prints one JSON object.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gen_protein_db(torch, device, n_seqs, seed):
    """-> (uint8 residues on host, uint64 offsets); BLOSUM62 background letters, lengths log-normal(ln 300, 0.6) in [30, 5000]"""
    import bench
    from mmseqs2_b200 import synth
    mat, pb = bench.load_matrix()
    cdf = torch.tensor(np.cumsum(synth.background(pb)), device=device, dtype=torch.float32)
    g = torch.Generator(device=device); g.manual_seed(seed)
    lens = torch.exp(np.log(300.0) + 0.6 * torch.randn(n_seqs, device=device, generator=g)).round().clamp(30, 5000).to(torch.int64)
    off = np.zeros(n_seqs + 1, np.uint64)
    off[1:] = torch.cumsum(lens, 0).cpu().numpy()
    total = int(off[-1])
    res = np.empty(total, np.uint8)
    chunk = 1 << 28
    for a in range(0, total, chunk):
        b = min(total, a + chunk)
        u = torch.rand(b - a, device=device, generator=g)
        res[a:b] = torch.clamp(torch.searchsorted(cdf, u, right=True), max=19).to(torch.uint8).cpu().numpy()
    return res, off


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--db3", type=int, default=20000000)
    ap.add_argument("--q3", type=int, default=48, help="config[3]: queries per rank (subsample of 100k / N)")
    ap.add_argument("--genome", type=float, default=5e9)
    ap.add_argument("--reads", type=int, default=1000000, help="config[4]: total reads over all ranks")
    ap.add_argument("--skip3", action="store_true")
    ap.add_argument("--skip4", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "multi_gpu_configs.json"))
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    import bench
    from mmseqs2_b200 import Context, SubMatrix, synth, alignment as al
    from mmseqs2_b200.sharding import HitGather
    mat, pb = bench.load_matrix()
    sm = SubMatrix(mat, pb)
    ctx = Context(local)
    out = {"n_gpus": world}

    def allsum(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def barrier():
        ctx.sync(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    # ---------------------------------------------------------------------------------------------------------------- config[3]
    if not args.skip3:
        t0 = time.perf_counter()
        res, off = gen_protein_db(torch, device, args.db3, seed=2)
        t_gen = time.perf_counter() - t0
        db_res = int(off[-1])
        rng = np.random.default_rng(1000 + rank)
        qres, qoff = synth.random_seqs(rng, args.q3, synth.background(pb), mean=350.0, sigma=35.0, lo=200, hi=500, normal=True)
        qs = synth.split(qres, qoff)
        # a few planted homologs so that the align step has survivors
        synth.plant_homologs(np.random.default_rng(3), res, off, qs[:8], synth.background(pb), frac=2000.0 / args.db3)
        t0 = time.perf_counter()
        ctx.load_db(res, off, 21)
        t_load = time.perf_counter() - t0
        profs = [sm.ssw_query(q) for q in qs]
        B = 16
        batches = [profs[i:i + B] for i in range(0, len(profs), B)]
        jobs = [ctx.scan_job(b, 15, 300) for b in batches]
        gather = HitGather(B, 300, dist, device) if world > 1 else None
        jobs[0].run(); ctx.sync()                       # warm-up
        barrier()
        t0 = time.perf_counter()
        depth = 2
        for s in range(min(depth, len(jobs))):
            jobs[s].run()
        lists = []
        for s, j in enumerate(jobs):
            h, nh, _ = j.fetch()
            if s + depth < len(jobs):
                jobs[s + depth].run()
            lists += [h[i]["id"][:int(nh[i])].copy() for i in range(len(h))]
            if gather is not None:
                if len(gather.started) >= 2:
                    gather.finish()
                gather.start(h, nh)
        if gather is not None:
            gather.drain()
        barrier()
        t_scan = time.perf_counter() - t0
        cells = float(sum(len(q) for q in qs)) * db_res
        evp = al.EvalueParams.defaults("blosum62.out", 11, 1, db_res)
        apar = al.AlignParams(sw_mode=al.SCORE_COV_SEQID, eval_thr=1e-3)
        al.align_batch(ctx, sm, qs[:4], lists[:4], apar, evp)
        barrier()
        t0 = time.perf_counter()
        ares, apool, n_aln = al.align_batch(ctx, sm, qs, lists, apar, evp)
        barrier()
        t_align = time.perf_counter() - t0
        tot_cells, t_scan_max, t_align_max = allsum(cells), allmax(t_scan), allmax(t_align)
        out["config3"] = {
            "workload": "easy-search shape: %d queries per GPU (L~350; subsample of the 100k-query job) vs a %d-sequence synthetic DB (%d residues, %.1f GB "
                        "resident per GPU), DB replicated, queries sharded x%d; ungapped prefilter -> top-300 lists (NCCL all_gather per 16-query batch) -> align -a -e 1e-3"
                        % (args.q3, args.db3, db_res, db_res / 1e9, world),
            "prefilter_GCUPS": tot_cells / 1e9 / t_scan_max, "prefilter_s": t_scan_max, "queries_total": int(allsum(len(qs))),
            "align_s": t_align_max, "alignments": int(allsum(n_aln)), "accepted": int(allsum(sum(len(r) for r in ares))),
            "queries_per_s_end_to_end": allsum(len(qs)) / (t_scan_max + t_align_max),
            "extrapolated_100k_queries_s": 100000.0 / (allsum(len(qs)) / (t_scan_max + t_align_max)),
            "db_generation_s": t_gen, "db_load_s": t_load}
        for j in jobs:
            j.close()
        del res, off
        if rank == 0:      # keep what is done even if a later part fails
            os.makedirs(os.path.dirname(args.out), exist_ok=True)
            json.dump(out, open(args.out, "w"), indent=1)
            print(json.dumps(out), flush=True)

    # ---------------------------------------------------------------------------------------------------------------- config[4]
    if not args.skip4:
        t0 = time.perf_counter()
        tl = 32000          # b200_nucl_align takes sequences up to 32767 (longer ones are split upstream, blastn.sh:27-52)
        n_t = int(args.genome // tl)
        g = torch.Generator(device=device); g.manual_seed(5)
        genome = np.empty(n_t * tl, np.uint8)
        chunk = 1 << 28
        for a in range(0, len(genome), chunk):
            b = min(len(genome), a + chunk)
            genome[a:b] = torch.randint(0, 4, (b - a,), device=device, generator=g, dtype=torch.uint8).cpu().numpy()
        goff = (np.arange(n_t + 1, dtype=np.uint64) * np.uint64(tl))
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        ctx.load_db(genome, goff, 5)
        t_load = time.perf_counter() - t0
        n_reads = args.reads // world
        rng = np.random.default_rng(40 + rank)
        tid = rng.integers(0, n_t, n_reads)
        pos = rng.integers(0, tl - 150, n_reads)
        comp = np.array([2, 3, 0, 1, 4], np.uint8)          # A,C,T,G,X = 0..4 (NucleotideMatrix.cpp): complement A<->T, C<->G
        reads, n_rev = [], 0
        for i in range(n_reads):
            w = genome[int(tid[i]) * tl + int(pos[i]):int(tid[i]) * tl + int(pos[i]) + 150]
            r = synth.nucl_mutate(rng, w, 0.02, 0.002)
            if i & 1:   # the read comes from the reverse strand: the aligner gets its reverse complement (BandedNucleotideAligner::initQuery)
                r = comp[comp[r[::-1]][::-1]]
                n_rev += 1
            reads.append(r)
        tasks = np.stack([np.arange(n_reads), tid, (-pos) & 0xffff], 1).astype(np.int64)
        packed = synth.pack(reads)
        ctx.nucl_align(reads[:2000], tasks[:2000], decode=False)
        barrier()
        t0 = time.perf_counter()
        nout, ncig, nbt = ctx.nucl_align(packed, tasks, decode=False)
        barrier()
        t_al = time.perf_counter() - t0
        k_ms = ctx.last_kernel_ms
        t_max, k_max = allmax(t_al), allmax(k_ms)
        tot = allsum(n_reads)
        out["config4"] = {
            "workload": "nucleotide mode: %d reads x 150 bp (2 %% subst, 0.2 %% indel, half from the reverse strand) vs a %.2g bp synthetic genome in %d targets of %d bp, "
                        "genome replicated (%.1f GB per GPU), reads sharded x%d; gapped aligner on the prefilter diagonal" % (int(tot), n_t * tl, n_t, tl, n_t * tl / 1e9, world),
            "alignments_per_s_e2e": tot / t_max, "alignments_per_s_kernel": tot / (k_max / 1e3), "e2e_s": t_max, "kernel_ms": k_max,
            "aligned_residues_per_s": allsum(float((nout["qend"] - nout["qstart"] + 1).sum())) / t_max,
            "mean_score": allsum(float(nout["score"].sum())) / tot, "reverse_strand_reads": int(allsum(n_rev)),
            "roofline": {"bound": "hbm", "bytes_per_read": 2 * (150 + 64) + 32, "achieved_gbs": tot * (2 * (150 + 64) + 32) / 1e9 / (k_max / 1e3)},
            "genome_generation_s": t_gen, "db_load_s": t_load}

    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
