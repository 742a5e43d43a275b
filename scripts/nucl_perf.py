"""nucleotide aligner throughput for the occupancy variants of nucl_align_kernel (B200_NUCL_MINB = 2, 3, 4), one process each"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np
    from mmseqs2_b200 import Context, synth
    rng = np.random.default_rng(4)
    targets = [synth.nucl_genome(rng, 30000) for _ in range(200)]
    reads, tasks = synth.nucl_reads(rng, targets, 200000, 150, subst=0.02, indel=0.002)
    td, to = synth.pack(targets)
    ctx = Context(0)
    ctx.load_db(td, to, 5)
    packed = synth.pack(reads)
    ctx.nucl_align(reads[:2000], tasks[:2000], decode=False)
    best = 1e9
    for _ in range(3):
        out, cig, bt = ctx.nucl_align(packed, tasks, decode=False)
        best = min(best, ctx.last_kernel_ms)
    print(json.dumps({"minb": os.environ.get("B200_NUCL_MINB", "2"), "kernel_ms": best, "alignments_per_s": len(reads) / (best / 1e3),
                      "checksum": int(out["score"].astype(np.int64).sum())}))
else:
    for mb in ("2", "3", "4"):
        env = dict(os.environ, B200_NUCL_MINB=mb)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-500:])
