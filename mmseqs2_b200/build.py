"""mmseqs2_b200/build.py -- compile libb200align.so in-tree (nvcc, sm_100a only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libb200align.so")
SOURCES = [os.path.join(HERE, "csrc", "b200_align.cu"), os.path.join(HERE, "csrc", "b200_nucl.cu"), os.path.join(HERE, "csrc", "b200_backtrace.cu"), os.path.join(HERE, "csrc", "b200_rescore.cu"), os.path.join(HERE, "csrc", "b200_host.cpp"),
           os.path.join(HERE, "csrc", "b200_alignment.cpp"), os.path.join(HERE, "csrc", "b200_db.cpp"), os.path.join(HERE, "csrc", "b200_gpuserver.cpp"), os.path.join(HERE, "csrc", "b200_multi.cpp"), os.path.join(HERE, "csrc", "b200_paddeddb.cpp"), os.path.join(HERE, "csrc", "b200_rescore_module.cpp")]
HEADERS = [os.path.join(ROOT, "include", "b200_align.h"), os.path.join(ROOT, "include", "b200_host.h"), os.path.join(HERE, "csrc", "b200_internal.h"),
           os.path.join(ROOT, "include", "b200_alignment.h"), os.path.join(ROOT, "include", "b200_db.h"), os.path.join(ROOT, "include", "b200_gpuserver.h"), os.path.join(ROOT, "include", "b200_multi.h")]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-I" + os.path.join(ROOT, "include"),
           "-o", SO] + SOURCES + ["-lrt", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
