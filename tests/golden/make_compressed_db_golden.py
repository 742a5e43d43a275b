#!/usr/bin/env python
"""tests/golden/make_compressed_db_golden.py -- a zstd-compressed DB triple written by the reference's own DBWriter
(WRITER_COMPRESSED_MODE, src/commons/DBWriter.cpp:372-413) through oracle/_ref, committed as bytes so that the reader test runs
where /root/reference does not exist.  Entries mix lengths below 60 bytes (stored raw, marker 0xFF), long compressible text, long
incompressible text and an empty entry.

  python tests/golden/make_compressed_db_golden.py        (needs oracle/_ref/libmmseqs_ref.so)
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def entries(rng):
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    out = [b"", b"M\n", bytes(aa[rng.integers(0, 20, 58)]) + b"\n", bytes(aa[rng.integers(0, 20, 59)]) + b"\n"]
    for n in (60, 61, 200, 1000, 5000):
        out.append(bytes(aa[rng.integers(0, 20, n)]) + b"\n")                         # protein text
    out.append(b"ACGT" * 2000 + b"\n")                                              # highly compressible
    out.append(bytes(rng.integers(1, 256, 3000).astype(np.uint8)))                    # incompressible, no NUL
    out.append(b"7\t120\t0.95\t1e-30\t0\t99\t100\t3\t102\t110\n" * 40)                # result-DB shaped text
    return out


def main():
    from oracle.pyoracle import Ref
    ref = Ref()
    rng = np.random.default_rng(77)
    ents = entries(rng)
    keys = rng.permutation(1000)[:len(ents)].astype(np.uint32)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "c")
        ref.db_write(path, 0, keys, ents, compressed=True)
        files = {suf: np.frombuffer(open(path + suf, "rb").read(), np.uint8) for suf in ("", ".index", ".dbtype")}
        k, l, o, ty = ref.db_read(path)                     # the reference's own DBReader on what it wrote
        order = np.argsort(keys, kind="stable")
        assert list(k) == [int(keys[i]) for i in order] and o == [ents[i] for i in order]
    width = max(len(e) for e in ents)
    mat = np.zeros((len(ents), width), np.uint8)
    for i, e in enumerate(ents):
        mat[i, :len(e)] = np.frombuffer(e, np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "compressed_db_v1.npz"), keys=keys, entries=mat,
                        entry_len=np.array([len(e) for e in ents], np.int64), file_data=files[""], file_index=files[".index"],
                        file_dbtype=files[".dbtype"], reader_dbtype=np.int64(ty))
    print("wrote compressed_db_v1.npz: %d entries, data %d bytes (plain %d)" % (len(ents), len(files[""]), sum(len(e) + 1 for e in ents)))


if __name__ == "__main__":
    main()
