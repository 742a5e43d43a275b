#!/usr/bin/env python
"""integration/search_wallclock.py -- `mmseqs search` wall-clock, B200 drop-in vs the reference's AVX2 CPU path (north_star target).

Both arms run the reference's own `search` workflow (src/workflow/Search.cpp, data/workflow/blastp.sh) on the same query DB and
the same padded target DB, at identical sensitivity (exhaustive ungapped prefilter, no k-mer stage):

  cpu   integration/_build/mmseqs_avx2 search Q T_pad res tmp --prefilter-mode 1 [-a] --threads N
            ungappedprefilter = runFilterOnCpu (ungappedprefilter.cpp:346-482), align = Alignment::run with per-thread Matchers
  b200  integration/_build/mmseqs_b200 search Q T_pad res tmp --gpu 1 [-a] --threads N
            ungappedprefilter = runFilterOnGpu (:41-343) with class Marv served by libb200align.so (integration/shim/marv.h),
            align = b200_align_batch in buckets (integration/shim/b200_align_module.h)

and the result DBs are compared: prefilter DB and alignment DB entry by entry (bytes), plus the md5 of the convertalis .m8.
Synthetic data as bench.py's config[1] (same generator, FASTA written here); `--examples DIR` uses the reference's example FASTA
files instead (BASELINE config[0]).  The CPU arm may run on a query subsample (--cpu-queries) and is then scaled linearly in queries
(the ungapped prefilter is linear in queries by construction; stated in the output).
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX", np.uint8)   # the numeric order of blosum62.out's alphabet (BaseMatrix::num2aa)


def write_fasta(path, res, off, prefix):
    off = np.asarray(off, np.int64)
    n = len(off) - 1
    letters = AA[res]
    with open(path, "wb") as f:
        chunk = []
        for i in range(n):
            chunk.append(b">%s%d\n" % (prefix, i))
            chunk.append(letters[off[i]:off[i + 1]].tobytes())
            chunk.append(b"\n")
            if len(chunk) >= 30000:
                f.write(b"".join(chunk)); chunk = []
        f.write(b"".join(chunk))


def run(cmd, log, env=None):
    t0 = time.perf_counter()
    with open(log, "ab") as lf:
        lf.write(("\n$ " + " ".join(cmd) + "\n").encode())
        lf.flush()
        rc = subprocess.call(cmd, stdout=lf, stderr=subprocess.STDOUT, env=env)
    if rc != 0:
        sys.stderr.write(open(log, errors="replace").read()[-3000:])
        raise SystemExit("command failed (%d): %s" % (rc, " ".join(cmd)))
    return time.perf_counter() - t0


def read_db(path):
    """-> {key: bytes} of an MMseqs2 DB (data [.N parts] + .index)"""
    parts = []
    if os.path.exists(path):
        parts = [path]
    else:
        k = 0
        while os.path.exists("%s.%d" % (path, k)):
            parts.append("%s.%d" % (path, k)); k += 1
    data = b"".join(open(p, "rb").read() for p in parts)
    out = {}
    for line in open(path + ".index"):
        k, o, l = line.split()
        out[int(k)] = data[int(o):int(o) + int(l)]
    return out


def md5(path):
    """md5 of the sorted lines: the alignment DB is written by several threads, so convertalis emits queries in varying order"""
    return hashlib.md5(b"".join(sorted(open(path, "rb").read().splitlines(True)))).hexdigest()


def step_times(log):
    """wall-clock of the modules as the workflow logs them (lines 'Time for processing: 0h 0m 1s 234ms')"""
    out, cur = [], None
    for line in open(log, errors="replace"):
        s = line.strip()
        if s.startswith("ungappedprefilter ") or s.startswith("align ") or s.startswith("prefilter "):
            cur = s.split()[0]
        if s.startswith("Time for processing:") and cur:
            parts = s.split(":", 1)[1].split()
            t = 0.0
            for p in parts:
                if p.endswith("ms"): t += float(p[:-2]) / 1e3
                elif p.endswith("h"): t += 3600 * float(p[:-1])
                elif p.endswith("m"): t += 60 * float(p[:-1])
                elif p.endswith("s"): t += float(p[:-1])
            out.append((cur, t)); cur = None
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--db-seqs", type=int, default=1000000)
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--cpu-queries", type=int, default=256, help="CPU arm runs on the first N queries (0 = skip the CPU arm)")
    ap.add_argument("--threads", type=int, default=0, help="0 = what the box really offers: min(affinity mask, cgroup cpu.max quota)")
    ap.add_argument("--work", default="/tmp/b200_search")
    ap.add_argument("--examples", default=None, help="directory with QUERY.fasta and DB.fasta (reference examples/)")
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--backtrace", action="store_true", help="search -a")
    ap.add_argument("--alignment-mode", type=int, default=None)
    ap.add_argument("--target-queries", type=int, default=100000, help="job size the two arms are extrapolated to (north_star: 100k x 1M)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "search_wallclock.json"))
    args = ap.parse_args()

    if args.threads <= 0:
        import bench as _b
        args.threads = _b.cpu_info()["threads_used"]
    cpu_bin = os.path.join(HERE, "_build", "mmseqs_avx2")
    gpu_bin = os.path.join(HERE, "_build", "mmseqs_b200")
    W = args.work
    shutil.rmtree(W, ignore_errors=True)
    os.makedirs(W)
    log = os.path.join(W, "log.txt")
    info = {"threads": args.threads, "affinity_cpus": len(os.sched_getaffinity(0)), "os_cpu_count": os.cpu_count()}
    try:
        info["cpu_model"] = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
        info["cgroup_cpu_max"] = open("/sys/fs/cgroup/cpu.max").read().strip()
    except Exception:
        pass

    # ---- data -----------------------------------------------------------------------------------------------------------
    t0 = time.perf_counter()
    if args.examples:
        qf, tf = os.path.join(args.examples, "QUERY.fasta"), os.path.join(args.examples, "DB.fasta")
        info["workload"] = "reference examples/QUERY.fasta vs examples/DB.fasta (BASELINE config[0])"
    else:
        import bench
        res, off, queries = bench.make_scan_workload(0, args.db_seqs, args.queries, 1)
        from mmseqs2_b200 import synth
        qres, qoff = synth.pack(queries)
        qf, tf = os.path.join(W, "q.fasta"), os.path.join(W, "t.fasta")
        write_fasta(qf, qres, qoff, b"q")
        write_fasta(tf, res, off, b"t")
        info["workload"] = "%d synthetic L~350 queries vs %d-sequence synthetic DB (bench.py config[1] generator, %d residues)" % (
            args.queries, args.db_seqs, int(off[-1]))
        info["cells"] = float(sum(len(q) for q in queries)) * float(off[-1])
    info["data_s"] = time.perf_counter() - t0
    thr = ["--threads", str(args.threads)]
    Q, T, TP = os.path.join(W, "Q"), os.path.join(W, "T"), os.path.join(W, "T_pad")
    info["createdb_s"] = run([cpu_bin, "createdb", qf, Q, "-v", "1"], log) + run([cpu_bin, "createdb", tf, T, "-v", "1"], log)
    info["makepaddedseqdb_s"] = run([cpu_bin, "makepaddedseqdb", T, TP, "-v", "1"] + thr, log)
    nq_total = len(open(Q + ".index").read().splitlines())
    info["queries"] = nq_total

    extra = []
    if args.backtrace:
        extra += ["-a"]
    if args.alignment_mode is not None:
        extra += ["--alignment-mode", str(args.alignment_mode)]
    out = {"info": info, "search_args": extra}

    def search(binary, qdb, name, mode_args):
        resdb, tmp = os.path.join(W, "res_" + name), os.path.join(W, "tmp_" + name)
        slog = os.path.join(W, "search_%s.log" % name)
        dt = run([binary, "search", qdb, TP, resdb, tmp] + mode_args + extra + thr + ["-v", "3"], slog)
        m8 = os.path.join(W, name + ".m8")
        run([cpu_bin, "convertalis", qdb, TP, resdb, m8, "-v", "1"] + thr, log)
        pref = os.path.join(tmp, "latest", "pref_0")
        return {"wall_s": dt, "modules": step_times(slog), "m8_md5": md5(m8), "m8_lines": sum(1 for _ in open(m8))}, resdb, pref

    # ---- B200 arm: all queries --------------------------------------------------------------------------------------------
    if not args.no_gpu:
        search(gpu_bin, Q, "b200_warm", ["--gpu", "1"])     # first run pays CUDA context creation + page-in of the binaries
        out["b200"], res_b, pref_b = search(gpu_bin, Q, "b200", ["--gpu", "1"])
        out["b200"]["queries"] = nq_total

    # ---- CPU arm: the first cpu_queries queries ----------------------------------------------------------------------------
    if args.cpu_queries:
        ncpu = min(args.cpu_queries, nq_total)
        QC = Q
        if ncpu < nq_total:
            QC = os.path.join(W, "Qsub")
            keys = os.path.join(W, "sub.keys")
            idx = sorted(int(l.split()[0]) for l in open(Q + ".index"))[:ncpu]
            open(keys, "w").write("".join("%d\n" % k for k in idx))
            run([cpu_bin, "createsubdb", keys, Q, QC, "--subdb-mode", "0", "-v", "1"], log)
        out["cpu"], res_c, pref_c = search(cpu_bin, QC, "cpu", ["--prefilter-mode", "1"])
        out["cpu"]["queries"] = ncpu
        out["cpu"]["wall_s_scaled_to_all_queries"] = out["cpu"]["wall_s"] * nq_total / ncpu
        if not args.no_gpu:
            # identical results on the shared queries: prefilter DB and alignment DB, entry by entry
            a, b = read_db(res_c), read_db(res_b)
            same_aln = all(a[k] == b.get(k) for k in a)
            pa, pb = read_db(pref_c), read_db(pref_b)
            same_pref = all(pa[k] == pb.get(k) for k in pa)
            out["parity"] = {"alignment_db_entries_compared": len(a), "alignment_db_identical": bool(same_aln),
                             "prefilter_db_entries_compared": len(pa), "prefilter_db_identical": bool(same_pref)}
            if not same_aln:
                bad = [k for k in a if a[k] != b.get(k)][:3]
                out["parity"]["first_differences"] = [{"key": k, "cpu": a[k][:300].decode(errors="replace"), "b200": (b.get(k) or b"")[:300].decode(errors="replace")} for k in bad]
            if ncpu == nq_total:
                out["parity"]["m8_md5_equal"] = out["cpu"]["m8_md5"] == out["b200"]["m8_md5"]
            out["speedup_search_wallclock"] = out["cpu"]["wall_s_scaled_to_all_queries"] / out["b200"]["wall_s"]
            if ncpu < nq_total and args.target_queries > 0:
                # Both arms are (fixed cost) + (per-query cost) x queries: start-up, DB load and index reading do not grow with the
                # query count, the ungapped prefilter and the alignments do.  The B200 arm is measured at two sizes (the CPU arm's
                # subsample and all queries), which gives its fixed and per-query parts; the CPU arm's fixed part is taken as zero
                # (the assumption that favours the CPU).  Extrapolation to the north_star job size, stated as such.
                out["b200_sub"], _, _ = search(gpu_bin, QC, "b200_sub", ["--gpu", "1"])
                per_q = (out["b200"]["wall_s"] - out["b200_sub"]["wall_s"]) / (nq_total - ncpu)
                fixed = out["b200"]["wall_s"] - per_q * nq_total
                cpu_per_q = out["cpu"]["wall_s"] / ncpu
                tq = args.target_queries
                out["extrapolation"] = {"target_queries": tq, "b200_fixed_s": fixed, "b200_per_query_s": per_q, "cpu_per_query_s": cpu_per_q,
                                        "b200_s": fixed + per_q * tq, "cpu_s": cpu_per_q * tq, "speedup": cpu_per_q * tq / (fixed + per_q * tq),
                                        "note": "linear in queries; CPU fixed cost taken as 0; same box, same thread count for both arms"}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
