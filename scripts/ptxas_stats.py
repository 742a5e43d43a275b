"""registers / spills of every ungapped_scan_kernel<G,K,TILED> instantiation (ptxas -v), for kernel tuning on the CPU box"""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
extra = sys.argv[1:]
out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xptxas", "-v", "-I" + os.path.join(ROOT, "include"),
                      "-c", os.path.join(ROOT, "mmseqs2_b200", "csrc", "b200_align.cu"), "-o", "/tmp/ptxas_stats.o"] + extra, capture_output=True, text=True).stderr
rows, cur = [], None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\w+)'", line)
    if m:
        cur = m.group(1)
    m2 = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m2 and cur:
        spill = (int(m2.group(2)), int(m2.group(3)))
    m3 = re.search(r"Used (\d+) registers", line)
    if m3 and cur:
        k = re.search(r"ungapped_scan_kernelILi(\d+)ELi(\d+)ELb([01])", cur)
        if k:
            rows.append((int(k.group(1)), int(k.group(2)), int(k.group(3)), int(m3.group(1)), spill))
        cur = None
for r in sorted(rows):
    print("G=%2d K=%2d tiled=%d regs=%3d spill st/ld=%s" % r)
