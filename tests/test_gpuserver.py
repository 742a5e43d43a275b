"""The `gpuserver` shared-memory protocol (include/b200_gpuserver.h): a B200 scan server answering the requests of a client that
follows the reference's state machine -- here the reference's own GPUSharedMemory client code (oracle/_ref) where it is built, and
a plain mmap client that walks the same header offsets."""
import ctypes
import mmap
import os
import struct
import threading
import time

import numpy as np
import pytest

from mmseqs2_b200.api import _p, _vp, load_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IDLE, RESERVED, READY, DONE = 0, 1, 2, 3


def mmap_client(name, q, profile):
    """GPUSharedMemory header: maxSeqLen, maxResListLen, state, serverExit, queryOffset, queryLen, resultsOffset, resultLen, profileOffset"""
    fd = os.open("/dev/shm" + name, os.O_RDWR)
    m = mmap.mmap(fd, 0)
    os.close(fd)
    hdr = struct.unpack_from("<IIiB3xIIIII", m, 0)
    q_off, r_off, p_off = hdr[4], hdr[6], hdr[8]
    while struct.unpack_from("<i", m, 8)[0] != IDLE:
        time.sleep(0)
    struct.pack_into("<i", m, 8, RESERVED)
    m[q_off:q_off + len(q)] = bytes(q)
    pb = np.ascontiguousarray(profile, np.int8).tobytes()
    m[p_off:p_off + len(pb)] = pb
    struct.pack_into("<I", m, 20, len(q))
    struct.pack_into("<i", m, 8, READY)
    while struct.unpack_from("<i", m, 8)[0] != DONE:
        assert struct.unpack_from("<B", m, 12)[0] == 0, "server exited"
        time.sleep(0)
    n = struct.unpack_from("<I", m, 28)[0]
    rec = np.frombuffer(m[r_off:r_off + 16 * n], np.int32).reshape(n, 4).copy()
    struct.pack_into("<i", m, 8, IDLE)
    m.close()
    return rec[:, 0].astype(np.uint32), rec[:, 1]


@pytest.mark.gpu
def test_server_answers_reference_protocol(ctx, submat):
    lib = load_library()
    ex = np.load(os.path.join(ROOT, "tests", "golden", "examples_v1.npz"))
    ctx.load_db(ex["tdata"], ex["toff"].astype(np.uint64), 21)
    qs = [ex["qdata"][int(ex["qoff"][i]):int(ex["qoff"][i + 1])] for i in range(12)]
    name = "/b200_gpuserver_test_%d" % os.getpid()
    srv = _vp()
    mat = np.ascontiguousarray(submat.mat, np.int16)
    rc = lib.b200_gpuserver_create(ctx.h, name.encode(), 4096, 300, _p(mat), 21, 15, ctypes.byref(srv))
    ctx._check(rc)
    lib.b200_gpuserver_served.restype = ctypes.c_uint64
    t = threading.Thread(target=lambda: lib.b200_gpuserver_serve(srv, ctypes.c_uint64(0)))
    t.start()
    try:
        from oracle.pyoracle import Ref
        ref = Ref() if Ref.available() else None
        n_req = 0
        for qi, q in enumerate(qs):
            prof = submat.ssw_query(q, comp_bias=(qi % 3 != 0))          # with and without composition bias
            exp = ex["ungapped"][qi].astype(np.int32) if qi % 3 != 0 else None
            hits, nh, dense = ctx_scan(ctx, prof)
            for client in ([mmap_client] + ([lambda n, a, b: ref.gpuserver_query(n, a, b)] if ref else [])):
                ids, sc = client(name, q, prof.profile)
                n_req += 1
                assert np.array_equal(ids, hits["id"][:nh]) and np.array_equal(sc, hits["score"][:nh]), qi
            if exp is not None:                                          # and the hits are the reference scorer's: > 15, (score desc, id asc), top 300
                sel = np.nonzero(exp > 15)[0]
                order = sel[np.lexsort((sel, -exp[sel]))][:300]
                assert np.array_equal(ids, order.astype(np.uint32)) and np.array_equal(sc, exp[order])
        assert int(lib.b200_gpuserver_served(srv)) == n_req
    finally:
        lib.b200_gpuserver_stop(srv)
        t.join(timeout=30)
        lib.b200_gpuserver_destroy(srv)
    assert not os.path.exists("/dev/shm" + name)


def ctx_scan(ctx, prof):
    hits, n_hits, dense = ctx.ungapped_scan([prof], min_score_excl=15, max_hits=300)
    return hits[0], int(n_hits[0]), dense
