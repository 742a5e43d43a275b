"""b200::DiagSubmitQueue (include/b200_mmseqs.hpp; SURVEY 8b seam B2: many host threads -> one device context).  The combining logic
is host code and is tested here with a stand-in backend (plain and under ThreadSanitizer); on a GPU the queue drives the real
b200_diag_score_batch from eight threads and must agree with direct b200_diag_score calls."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = "-I" + os.path.join(ROOT, "include")


def _build(tmp_path, name, extra=()):
    exe = str(tmp_path / name)
    subprocess.check_call(["g++", "-std=c++11", "-fno-exceptions", "-O1", "-Wall", "-pthread", INC, os.path.join(ROOT, "tests", "cpp", "submit_queue_test.cpp"),
                           "-o", exe] + list(extra))
    return exe


def test_queue_combines_and_returns_each_thread_its_own_results(tmp_path):
    exe = _build(tmp_path, "sq")
    for args in (["16", "60"], ["1", "50"], ["3", "100"], ["32", "40", "5"]):       # last: every 5th backend call fails -> its whole round sees the error
        out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
    fields = dict(zip(out.stdout.split()[0::2], out.stdout.split()[1::2]))
    assert int(fields["failed"]) > 0 and int(fields["wrong"]) == 0


def test_queue_under_thread_sanitizer(tmp_path):
    try:
        exe = _build(tmp_path, "sq_tsan", ["-g", "-fsanitize=thread"])
    except subprocess.CalledProcessError:
        pytest.skip("no ThreadSanitizer runtime here")
    out = subprocess.run([exe, "12", "30"], capture_output=True, text=True, timeout=300)
    if "FATAL: ThreadSanitizer" in out.stderr and "unexpected memory mapping" in out.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory on this kernel")
    assert out.returncode == 0 and "WARNING: ThreadSanitizer" not in out.stderr, out.stdout + out.stderr


@pytest.mark.gpu
def test_queue_on_the_device_equals_direct_calls(tmp_path, built_lib):
    exe = str(tmp_path / "sq_gpu")
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["g++", "-std=c++11", "-fno-exceptions", "-O1", "-pthread", INC, os.path.join(ROOT, "tests", "cpp", "submit_queue_gpu.cpp"),
                           "-L" + libdir, "-lb200align", "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
